// lbm-lid-driven-cavity-flow — the reference benchmark's flow on the B200-native kernel library.
//
// Same command line, same problem set-up, same metric and the same report keys as
// benchmarks/lbm-lid-driven-cavity-flow (src/app.cpp, src/Config.cpp:58-111, src/RunCavityTwoPop.cu:20-317,
// src/Metrics.h:32-50, src/Report.cpp:5-102), written against the C++ veneer (Neon/…) over libneon_lbm.so:
//
//   lbm-lid-driven-cavity-flow --deviceType gpu --deviceIds 0 [1 ...] --grid dGrid --domain-size N
//        --warmup-iter W --max-iter M --repetitions R --report-filename F --computeFP float|double --storageFP float|double
//        [--sOCC|--nOCC] [--put|--get] [--huLattice|--huGrid] [--benchmark|--visual] [--vti]
// Extensions (not in the reference): --lattice D3Q19|D3Q27, --arith fast|reference, --geom cavity|sphere,
//   --dim NX NY NZ (non-cubic box), --device-setup (problem set-up on the device), --graph (CUDA graph
//   replay, one device), --dump FILE (populations, wall masks and classes in oracle/ref_driver.cu's format).
//
// --deviceType cpu is refused: there is no CPU compute path (the reference's CPU numbers come from the reference).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "Neon/Neon.h"
#include "Neon/Report.h"
#include "Neon/domain/bGrid.h"
#include "Neon/domain/dGrid.h"
#include "Neon/lbm/Lbm.h"
#include "Neon/set/Backend.h"
#include "Neon/skeleton/Skeleton.h"

namespace {

struct LbmParameters
{
    double nu = 0, omega = 0, dx = 0, dt = 0;
};

struct Config
{
    double                     Re = 100.;
    double                     ulb = 0.04;
    int                        N = 160;
    bool                       benchmark = false;
    double                     max_t = 10.0;
    int                        outFrequency = 200;
    int                        dataFrequency = 0;
    int                        benchIniIter = 1000;
    int                        benchMaxIter = 2000;
    int                        repetitions = 1;
    std::string                deviceType = "gpu";
    std::vector<int>           devices;
    std::string                reportFile = "lbm-lid-driven-cavity-flow";
    std::string                gridType = "dGrid";
    Neon::skeleton::Occ        occ = Neon::skeleton::Occ::none;
    Neon::set::TransferMode    transferMode = Neon::set::TransferMode::get;
    Neon::set::StencilSemantic stencilSemantic = Neon::set::StencilSemantic::streaming;
    bool                       vti = false;
    std::string                computeType = "double";
    std::string                storeType = "double";
    LbmParameters              lbm;
    // extensions
    std::string lattice = "D3Q19";
    std::string arith = "fast";
    std::string geom = "cavity";
    std::string dump;
    int         dim[3] = {0, 0, 0};
    bool        deviceSetup = false;
    bool        cudaGraph = false;

    static void usage(const char* argv0)
    {
        std::cout << "SYNOPSIS\n  " << argv0
                  << " --deviceType <cpu|gpu> --deviceIds <id>... [--grid <dGrid|bGrid|eGrid>] [--domain-size <N>] [--warmup-iter <W>]\n"
                     "      [--max-iter <M>] [--repetitions <R>] [--report-filename <F>] [--computeFP <float|double>]\n"
                     "      [--storageFP <float|double>] [--sOCC|--nOCC] [--put|--get] [--huLattice|--huGrid] [--benchmark|--visual] [--vti]\n"
                     "      [--lattice <D3Q19|D3Q27>] [--arith <fast|reference>] [--geom <cavity|sphere>] [--dim <NX> <NY> <NZ>]\n"
                     "      [--device-setup] [--graph] [--dump <file>]\n";
    }

    int parseArgs(int argc, char* argv[])
    {
        bool haveType = false, haveIds = false;
        for (int i = 1; i < argc; ++i) {
            const std::string k = argv[i];
            if (k == "--help" || k == "-h") {
                usage(argv[0]);
                return 1;
            }
            auto value = [&]() -> std::string {
                if (i + 1 >= argc) {
                    throw std::runtime_error("missing value after " + k);
                }
                return argv[++i];
            };
            try {
                if (k == "--deviceType") {
                    deviceType = value();
                    haveType = true;
                } else if (k == "--deviceIds") {
                    while (i + 1 < argc && (std::isdigit(argv[i + 1][0]) != 0)) {
                        devices.push_back(std::atoi(argv[++i]));
                    }
                    haveIds = !devices.empty();
                } else if (k == "--grid") {
                    gridType = value();
                } else if (k == "--domain-size") {
                    N = std::stoi(value());
                } else if (k == "--warmup-iter") {
                    benchIniIter = std::stoi(value());
                } else if (k == "--max-iter") {
                    benchMaxIter = std::stoi(value());
                } else if (k == "--repetitions") {
                    repetitions = std::stoi(value());
                } else if (k == "--report-filename") {
                    reportFile = value();
                } else if (k == "--computeFP") {
                    computeType = value();
                } else if (k == "--storageFP") {
                    storeType = value();
                } else if (k == "--sOCC") {
                    occ = Neon::skeleton::Occ::standard;
                } else if (k == "--nOCC") {
                    occ = Neon::skeleton::Occ::none;
                } else if (k == "--put") {
                    transferMode = Neon::set::TransferMode::put;
                } else if (k == "--get") {
                    transferMode = Neon::set::TransferMode::get;
                } else if (k == "--huLattice") {
                    stencilSemantic = Neon::set::StencilSemantic::streaming;
                } else if (k == "--huGrid") {
                    stencilSemantic = Neon::set::StencilSemantic::standard;
                } else if (k == "--benchmark") {
                    benchmark = true;
                } else if (k == "--visual") {
                    benchmark = false;
                } else if (k == "--vti") {
                    vti = true;
                } else if (k == "--lattice") {
                    lattice = value();
                } else if (k == "--arith") {
                    arith = value();
                } else if (k == "--geom") {
                    geom = value();
                } else if (k == "--dump") {
                    dump = value();
                } else if (k == "--dim") {
                    for (int& d : dim) {
                        d = std::stoi(value());
                    }
                } else if (k == "--device-setup") {
                    deviceSetup = true;
                } else if (k == "--graph") {
                    cudaGraph = true;
                } else {
                    throw std::runtime_error("unknown option " + k);
                }
            } catch (const std::exception& e) {
                std::cout << e.what() << "\n";
                usage(argv[0]);
                return -1;
            }
        }
        if (!haveType || !haveIds) {
            usage(argv[0]);
            return -1;
        }
        if (dim[0] == 0) {
            dim[0] = dim[1] = dim[2] = N;
        } else {
            N = dim[0];
        }
        /* Config::helpSetLbmParameters, Config.cpp:105-111 */
        lbm.nu = ulb * static_cast<double>(N - 2) / Re;
        lbm.omega = 1. / (3. * lbm.nu + 0.5);
        lbm.dx = 1. / static_cast<double>(N - 2);
        lbm.dt = lbm.dx * ulb;
        return 0;
    }

    std::string toString() const
    {
        std::ostringstream s;
        std::string        ids;
        for (int d : devices) {
            ids += (ids.empty() ? "" : " ") + std::to_string(d);
        }
        s << ".................. Re " << Re << "\n................. ulb " << ulb << "\n................... N " << N
          << "\n........... benchmark " << benchmark << "\n........ benchIniIter " << benchIniIter << "\n........ benchMaxIter "
          << benchMaxIter << "\n.......... deviceType " << deviceType << "\n.......... numDevices " << devices.size()
          << "\n............. devices " << ids << "\n.......... reportFile " << reportFile << "\n............ gridType " << gridType
          << "\n......... computeType " << computeType << "\n........... storeType " << storeType << "\n. ............... occ "
          << Neon::skeleton::OccUtils::toString(occ) << "\n....... transfer Mode " << Neon::set::TransferModeUtils::toString(transferMode)
          << "\n... transfer Semantic " << Neon::set::StencilSemanticUtils::toString(stencilSemantic) << "\n............. lattice "
          << lattice << "\n............... arith " << arith << "\n. ............... nu " << lbm.nu << "\n.............. omega "
          << lbm.omega << "\n................. dx " << lbm.dx << "\n................. dt " << lbm.dt << "\n";
        return s.str();
    }
};

/* the benchmark's Report (src/Report.cpp): configuration keys up front, result vectors at save() */
struct RunReport
{
    Neon::Report        report{"lbm-lid-driven-cavity-flow"};
    std::string         fname;
    std::vector<double> mlups, loopTime, setupTime, gridInitTime;
    bool                backendRecorded = false;

    explicit RunReport(const Config& c) : fname(c.reportFile)
    {
        report.addMember("Re", c.Re);
        report.addMember("ulb", c.ulb);
        report.addMember("N", c.N);
        report.addMember("benchmark", c.benchmark);
        report.addMember("max_t", c.max_t);
        report.addMember("outFrequency", c.outFrequency);
        report.addMember("dataFrequency", c.dataFrequency);
        report.addMember("repetitions", c.repetitions);
        report.addMember("vti", c.vti);
        report.addMember("benchIniIter", c.benchIniIter);
        report.addMember("benchMaxIter", c.benchMaxIter);
        report.addMember("deviceType", c.deviceType);
        report.addMember("numDevices", c.devices.size());
        report.addMember("devices", c.devices);
        report.addMember("reportFile", c.reportFile);
        report.addMember("gridType", c.gridType);
        report.addMember("computeType", c.computeType);
        report.addMember("storeType", c.storeType);
        report.addMember("occ", Neon::skeleton::OccUtils::toString(c.occ));
        report.addMember("transferMode", Neon::set::TransferModeUtils::toString(c.transferMode));
        report.addMember("transferSemantic", Neon::set::StencilSemanticUtils::toString(c.stencilSemantic));
        report.addMember("nu", c.lbm.nu);
        report.addMember("omega", c.lbm.omega);
        report.addMember("dx", c.lbm.dx);
        report.addMember("dt", c.lbm.dt);
        report.addMember("lattice", c.lattice);
        report.addMember("arith", c.arith);
        report.addMember("dim", std::vector<int>{c.dim[0], c.dim[1], c.dim[2]});
    }
    void save()
    {
        report.addMember("MLUPS", mlups);
        report.addMember("Loop Time (microseconds)", loopTime);
        report.addMember("Problem Setup Time (microseconds)", setupTime);
        report.addMember("Neon Grid Init Time (microseconds)", gridInitTime);
        std::cout << "Report: " << report.write(fname, true) << std::endl;
    }
};

using Clock = std::chrono::high_resolution_clock;
double microsSince(const Neon::Backend& bk, Clock::time_point start)
{
    bk.syncAll();
    return double(std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - start).count());
}

bool inSphere(const Config& c, const Neon::index_3d& p)
{
    if (c.geom != "sphere") {
        return false;
    }
    /* the obstacle of oracle/ref_driver.cu: off-centre, radius min(dim)/5 */
    const double cx = 0.45 * c.dim[0], cy = 0.55 * c.dim[1], cz = 0.5 * c.dim[2];
    const double R = std::min({c.dim[0], c.dim[1], c.dim[2]}) / 5.0;
    const double dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
    return dx * dx + dy * dy + dz * dz < R * R;
}

template <typename Lattice, typename Grid, typename StorageFP, typename ComputeFP>
void run(Config& config, RunReport& report)
{
    using PopulationField = typename Grid::template Field<StorageFP, Lattice::Q>;
    using Tools = LbmContainers<Lattice, PopulationField, ComputeFP>;

    if (config.deviceType != "gpu") {
        Neon::NeonException e("run");
        e << "deviceType '" << config.deviceType << "': only gpu is supported — there is no CPU compute path behind this veneer";
        NEON_THROW(e);
    }
    Neon::Backend bk(config.devices, Neon::Runtime::stream);
    if (!report.backendRecorded) {
        bk.toReport(report.report);
        report.backendRecorded = true;
    }
    Neon::lbm::kernelOptions() = config.arith == "reference" ? NLBM_ARITH_REFERENCE : NLBM_ARITH_FAST;

    const Neon::double_3d ulid(1., 0., 0.);
    Lattice               lattice(bk);
    const Neon::index_3d  dim(config.dim[0], config.dim[1], config.dim[2]);

    auto start = Clock::now();
    Grid grid(bk, dim, [](const Neon::index_3d&) { return true; }, lattice.c_vect);
    PopulationField pop0 = grid.template newField<StorageFP, Lattice::Q>("Population", Lattice::Q, StorageFP(0.0));
    PopulationField pop1 = grid.template newField<StorageFP, Lattice::Q>("Population", Lattice::Q, StorageFP(0.0));
    typename Grid::template Field<StorageFP, 1> rho;
    typename Grid::template Field<StorageFP, 3> u;
    if (!config.benchmark) {
        rho = grid.template newField<StorageFP, 1>("rho", 1, StorageFP(0.0));
        u = grid.template newField<StorageFP, 3>("u", 3, StorageFP(0.0));
    }
    auto flag = grid.template newField<CellType, 1>("Material", 1, CellType());
    const ComputeFP omega = static_cast<ComputeFP>(config.lbm.omega);

    LbmIterationT<Lattice, PopulationField, ComputeFP> iteration(config.stencilSemantic, config.occ, config.transferMode, pop0, pop1,
                                                                  flag, omega, config.cudaGraph);
    report.gridInitTime.push_back(microsSince(bk, start));
    std::cout << "Metrics:\n    Grid Init: " << report.gridInitTime.back() << " microseconds" << std::endl;

    // ---- problem set-up (RunCavityTwoPop.cu:159-242): host loops over the mirrors, upload, halo, wall mask ----------
    start = Clock::now();
    auto isEdge = [&](const Neon::index_3d& p) {
        return p.x == 0 || p.x == dim.x - 1 || p.y == 0 || p.y == dim.y - 1 || p.z == 0 || p.z == dim.z - 1;
    };
    const auto&  t = lattice.t_vect;
    const auto&  c = lattice.c_vect;
    const double ulb = config.ulb;
    auto         initPop = [&](const Neon::index_3d& p, const int& k, StorageFP& val) {
        val = static_cast<StorageFP>(t.at(k));
        if (isEdge(p)) {
            if (p.y == dim.y - 1) {
                if constexpr (Lattice::Q == 19) {
                    val = static_cast<StorageFP>(-6. * t.at(k) * ulb * (c.at(k).v[0] * ulid.v[0] + c.at(k).v[1] * ulid.v[1] + c.at(k).v[2] * ulid.v[2]));
                } else { /* apps/lbmMultiRes/lidDrivenCavity.h:56-76: the dot product is accumulated in the storage type */
                    StorageFP dot = 0;
                    const double wall[3] = {ulb, 0., 0.};
                    for (int d = 0; d < 3; ++d) {
                        dot = static_cast<StorageFP>(double(dot) + double(c.at(k).v[d]) * wall[d]);
                    }
                    val = static_cast<StorageFP>(double(dot) * (-6. * t.at(k)));
                }
            } else {
                val = 0;
            }
        } else if (inSphere(config, p)) {
            val = 0;
        }
    };
    if (config.deviceSetup) {
        /* classes, wall masks and initial populations produced on the device */
        const int geomId = config.geom == "sphere" ? 1 : 0;
        for (int d = 0; d < bk.getDeviceCount(); ++d) {
            bk.setDevice(d);
            auto desc = pop0.getPartition(d).desc; /* nlbm_dense_desc or nlbm_block_desc */
            desc.flags = flag.getPartition(d).mem();
            cudaStream_t st = bk.stream(d, 0);
            if constexpr (std::is_same_v<Grid, Neon::bGrid>) {
                Neon::detail::check(nlbm_block_classify(&desc, geomId, nullptr, grid.activeMaskDev(d), st), "nlbm_block_classify");
                Neon::detail::check(nlbm_block_wall_mask(&desc, Lattice::Q, nullptr, st), "nlbm_block_wall_mask");
            } else {
                Neon::detail::check(nlbm_dense_classify(&desc, geomId, nullptr, st), "nlbm_dense_classify");
                Neon::detail::check(nlbm_dense_wall_mask(&desc, Lattice::Q, nullptr, st), "nlbm_dense_wall_mask");
            }
            for (auto* f : {&pop0, &pop1}) {
                desc.pop_out = f->getPartition(d).mem();
                if constexpr (std::is_same_v<Grid, Neon::bGrid>) {
                    if constexpr (std::is_same_v<StorageFP, float>) {
                        Neon::detail::check(nlbm_block_init_pop_f32(&desc, Lattice::Q, ulb, st), "nlbm_block_init_pop_f32");
                    } else {
                        Neon::detail::check(nlbm_block_init_pop_f64(&desc, Lattice::Q, ulb, st), "nlbm_block_init_pop_f64");
                    }
                } else if constexpr (std::is_same_v<StorageFP, float>) {
                    Neon::detail::check(nlbm_dense_init_pop_f32(&desc, Lattice::Q, ulb, st), "nlbm_dense_init_pop_f32");
                } else {
                    Neon::detail::check(nlbm_dense_init_pop_f64(&desc, Lattice::Q, ulb, st), "nlbm_dense_init_pop_f64");
                }
            }
        }
        pop0.commitWalls();
        pop1.commitWalls();
        bk.syncAll();
    } else {
        iteration.getInput().forEachActiveCell(initPop);
        iteration.getOutput().forEachActiveCell(initPop);
        flag.forEachActiveCell([&](const Neon::index_3d& p, const int&, CellType& f) {
            f.classification = CellType::bulk;
            f.wallNghBitflag = 0;
            if (isEdge(p)) {
                f.classification = p.y == dim.y - 1 ? CellType::movingWall : CellType::bounceBack;
            } else if (inSphere(config, p)) {
                f.classification = CellType::bounceBack;
            }
        });
        iteration.getInput().updateDeviceData(Neon::Backend::mainStreamIdx);
        iteration.getOutput().updateDeviceData(Neon::Backend::mainStreamIdx);
        flag.updateDeviceData(Neon::Backend::mainStreamIdx);
        bk.syncAll();
        flag.newHaloUpdate(Neon::set::StencilSemantic::standard, Neon::set::TransferMode::get, Neon::Execution::device)
            .run(Neon::Backend::mainStreamIdx);
        bk.syncAll();
        Tools::computeWallNghMask(flag, flag).run(Neon::Backend::mainStreamIdx);
        bk.syncAll();
    }
    report.setupTime.push_back(microsSince(bk, start));
    std::cout << "Metrics:\n    Problem Setup: " << report.setupTime.back() << " microseconds" << std::endl;

    // ---- --visual: rho/u export every 100 iterations (RunCavityTwoPop.cu:82-150) --------------------------------------
    // one halo-update container per population field, built on first use and kept (a fresh one every export would leave
    // its events with the Backend until the end of the run)
    std::map<size_t, Neon::set::Container> visualHalo;
    auto exportRhoAndU = [&](int iterationId) {
        if constexpr (Lattice::Q == 19 && std::is_same_v<Grid, Neon::dGrid>) {
            if (iterationId % 100 != 0) {
                return;
            }
            auto& f = iteration.getInput();
            bk.syncAll();
            if (!visualHalo.count(f.getUid())) {
                visualHalo.emplace(f.getUid(), f.newHaloUpdate(Neon::set::StencilSemantic::standard, Neon::set::TransferMode::get,
                                                               Neon::Execution::device));
            }
            visualHalo.at(f.getUid()).run(Neon::Backend::mainStreamIdx);
            bk.syncAll();
            Tools::computeRhoAndU(f, flag, rho, u).run(Neon::Backend::mainStreamIdx);
            u.updateHostData(Neon::Backend::mainStreamIdx);
            rho.updateHostData(Neon::Backend::mainStreamIdx);
            bk.syncAll();
            std::string id = std::to_string(iterationId);
            id = std::string(5 - std::min<size_t>(5, id.length()), '0') + id;
            u.ioToVtk("u_" + id, "u", false);
            rho.ioToVtk("rho_" + id, "rho", false);
            /* centre-line profiles: u_x(y) at x = z = N/2 and u_y(x) at y = z = N/2, scaled by the lid speed */
            std::ofstream fy("NeonUniformLBM_" + id + "_Y.dat"), fx("NeonUniformLBM_" + id + "_X.dat");
            const double  scale = 1.0 / ulid.v[0];
            for (int y = 0; y < dim.y; ++y) {
                fy << double(y) / double(dim.y) << " " << u(Neon::index_3d(dim.x / 2, y, dim.z / 2), 0) * scale << "\n";
            }
            for (int x = 0; x < dim.x; ++x) {
                fx << double(x) / double(dim.x) << " " << u(Neon::index_3d(x, dim.y / 2, dim.z / 2), 1) * scale << "\n";
            }
        }
    };

    // ---- time loop and metric (RunCavityTwoPop.cu:244-275, Metrics.h:32-50) --------------------------------------------
    start = Clock::now();
    int clockIter = 0;
    /* benchmark mode on one dense partition of a small box: up to 10 iterations per library call (the launch chain of
     * nlbm_dense_step_n), never across the end of the warm-up */
    const bool chain = config.benchmark && !config.cudaGraph && iteration.chainPays();
    for (int it = 0; it < config.benchMaxIter;) {
        if (!config.benchmark) {
            exportRhoAndU(it);
        }
        if (config.benchmark && it == config.benchIniIter) {
            std::cout << "Warm up completed (" << it << " iterations ).\nStarting benchmark step ("
                      << config.benchMaxIter - config.benchIniIter << " iterations)." << std::endl;
            bk.syncAll();
            start = Clock::now();
            clockIter = 0;
        }
        int n = 1;
        if (chain) {
            const int stop = it < config.benchIniIter ? config.benchIniIter : config.benchMaxIter;
            n = std::min(10, stop - it);
            iteration.runMany(n);
        } else {
            iteration.run();
        }
        it += n;
        clockIter += n;
    }
    std::cout << "Iterations completed" << std::endl;
    const double us = microsSince(bk, start);
    const double mlups = double(dim.rMul<size_t>()) * double(clockIter) / us;
    report.loopTime.push_back(us);
    report.mlups.push_back(mlups);
    std::cout << "Metrics: \n     time: " << std::setprecision(4) << us << " microseconds\n    MLUPS: " << std::setprecision(6) << mlups
              << " MLUPS" << std::endl;

    if (!config.dump.empty()) {
        auto& f = iteration.getInput();
        f.updateHostData(Neon::Backend::mainStreamIdx);
        flag.updateHostData(Neon::Backend::mainStreamIdx);
        bk.syncAll();
        const size_t          cells = dim.rMul<size_t>();
        std::vector<uint32_t> mask(cells);
        std::vector<int32_t>  cls(cells);
        flag.forEachActiveCell(
            [&](const Neon::index_3d& p, const int&, CellType& v) {
                const size_t o = (size_t(p.z) * dim.y + p.y) * dim.x + p.x;
                mask[o] = v.wallNghBitflag;
                cls[o] = static_cast<int32_t>(v.classification);
            },
            Neon::computeMode_t::seq);
        FILE* fp = std::fopen(config.dump.c_str(), "wb");
        if (!fp) {
            Neon::NeonException e("dump");
            e << "cannot open " << config.dump;
            NEON_THROW(e);
        }
        const int32_t hdr[8] = {0x4E4C424D, dim.x, dim.y, dim.z, Lattice::Q, int32_t(sizeof(StorageFP)), config.benchMaxIter,
                                config.geom == "sphere" ? 1 : 0};
        const double  om = config.lbm.omega;
        std::fwrite(hdr, sizeof(hdr), 1, fp);
        std::fwrite(&om, sizeof(double), 1, fp);
        std::fwrite(f.hostData(), sizeof(StorageFP), cells * Lattice::Q, fp); /* the mirror is [q][z][y][x] already */
        std::fwrite(mask.data(), sizeof(uint32_t), cells, fp);
        std::fwrite(cls.data(), sizeof(int32_t), cells, fp);
        std::fclose(fp);
    }
}

template <typename Lattice19, typename Lattice27, typename Grid, typename S, typename C>
void runLattice(Config& config, RunReport& report)
{
    if (config.lattice == "D3Q19") {
        return run<Lattice19, Grid, S, C>(config, report);
    }
    if constexpr (std::is_same_v<S, C>) {
        if (config.lattice == "D3Q27") {
            return run<Lattice27, Grid, S, C>(config, report);
        }
    }
    NEON_THROW_UNSUPPORTED_OPERATION("lattice " + config.lattice + " with this precision pair");
}

template <typename Grid>
void runPrecision(Config& config, RunReport& report)
{
    const std::string& s = config.storeType;
    const std::string& c = config.computeType;
    if (s == "double" && c == "double") {
        return runLattice<D3Q19Template<double, double>, D3Q27Template<double, double>, Grid, double, double>(config, report);
    }
    if (s == "float" && c == "double") {
        return runLattice<D3Q19Template<float, double>, D3Q27Template<float, double>, Grid, float, double>(config, report);
    }
    if (s == "float" && c == "float") {
        return runLattice<D3Q19Template<float, float>, D3Q27Template<float, float>, Grid, float, float>(config, report);
    }
    NEON_THROW_UNSUPPORTED_OPERATION("storageFP " + s + " with computeFP " + c);
}

}  // namespace

int main(int argc, char* argv[])
{
    Config    config;
    const int parsed = config.parseArgs(argc, argv);
    if (parsed != 0) {
        return parsed > 0 ? 0 : -1; /* --help: 0 */
    }
    try {
        Neon::init();
        std::cout << "--------------- Parameters ---------------\n" << config.toString() << "-------------------------------------------\n";
        RunReport report(config);
        for (int r = 0; r < config.repetitions; ++r) {
            if (config.gridType == "dGrid") {
                runPrecision<Neon::dGrid>(config, report);
            } else if (config.gridType == "bGrid") {
                runPrecision<Neon::bGrid>(config, report);
            } else if (config.gridType == "eGrid") {
                /* the reference's element-sparse grid: the cavity activates every cell, so the dense layout serves it
                 * (keeps the --grid axis of the reference sweep complete; recorded in the report) */
                if (r == 0) {
                    report.report.addMember("gridServedBy", std::string("dGrid (every cell of the box is active)"));
                }
                runPrecision<Neon::dGrid>(config, report);
            } else {
                NEON_THROW_UNSUPPORTED_OPERATION("grid " + config.gridType + " (dGrid and bGrid are on the accelerated path; eGrid is not)");
            }
        }
        report.save();
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
