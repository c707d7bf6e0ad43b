// TEST INFRASTRUCTURE — not part of the product path.
//
// Driver for the UNMODIFIED reference (Autodesk/Neon v0.3.3) LBM hot path on
// its CPU/OpenMP backend (default) or, with --device gpu, on its own CUDA
// backend (generic lambda kernel compiled for sm_100: the reference's GPU
// number on the same B200).  This file is our own code; it is compiled against
// the reference headers and sources where they lie under /root/reference (see
// oracle/Makefile.ref) and only into oracle/_ref/.  It exists because the
// stock benchmark never dumps populations
// (benchmarks/lbm-lid-driven-cavity-flow/src/RunCavityTwoPop.cu:98,107 are
// commented out), and population-level parity is what the north star asks.
//
// What it runs is the reference's own code, untouched:
//   LbmIterationD3Q19 (src/LbmIteration.h:19-101)  -> two Skeletons
//   LbmContainers::iteration (src/LbmTools.h:285-325)
//   LbmContainers::computeWallNghMask (src/LbmTools.h:344-376)
// The problem set-up follows RunCavityTwoPop.cu:159-242 (cavity) and adds an
// optional solid sphere (bounce-back cells inside the cavity) so that
// obstacle bounce-back is pinned by the real reference kernel as well.
//
// Output (--dump FILE): little-endian binary
//   int32 magic 0x4E4C424D, int32 nx, ny, nz, int32 q, int32 fp_bytes,
//   int32 iters, int32 geom, float64 omega,
//   populations  [q][z][y][x]  (fp_bytes each)  of the CURRENT input field,
//   wall masks   [z][y][x] uint32,
//   class        [z][y][x] int32.
// --bench W : times (iters-W) iterations after W warm-up ones and prints one
//   JSON line with MLUPS computed as in src/Metrics.h:39-42.

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "Neon/Neon.h"
#include "Neon/domain/bGrid.h"
#include "Neon/domain/dGrid.h"

#include "CellType.h"
#include "D3Q19.h"
#include "LbmIteration.h"

struct Args
{
    int         nx = 16, ny = 16, nz = 16;
    int         iters = 10;
    int         warmup = -1;
    int         nDev = 1;
    int         geom = 0;  // 0 cavity, 1 cavity + sphere obstacle
    bool        isDouble = false;
    std::string grid = "dGrid";
    std::string dump;
    bool        occ = false;      // --occ standard: the reference's Skeleton splits INTERNAL / BOUNDARY (Occ::standard)
    bool        sameGpu = false;  // --same-gpu: every partition on device 0 (oversubscribed device list, as the reference's tests do)
    std::string device = "cpu";  // "gpu": the reference's own CUDA backend (Neon::Runtime::stream), devices 0..nDev-1
    double      Re = 100., ulb = 0.04;
};

static bool isSolidSphere(const Args& a, int x, int y, int z)
{
    if (a.geom != 1)
        return false;
    // Sphere centred slightly off-centre so that no symmetry hides an
    // indexing mistake; radius = min(dim)/5.
    const double cx = 0.45 * a.nx, cy = 0.55 * a.ny, cz = 0.5 * a.nz;
    int          m = a.nx < a.ny ? a.nx : a.ny;
    m = m < a.nz ? m : a.nz;
    const double R = m / 5.0;
    const double dx = x - cx, dy = y - cy, dz = z - cz;
    return dx * dx + dy * dy + dz * dz < R * R;
}

template <typename Grid, typename FP>
static int runCase(const Args& a)
{
    using Lattice = D3Q19Template<FP, FP>;
    using PopulationField = typename Grid::template Field<FP, Lattice::Q>;

    std::vector<int> devs(a.nDev, 0);
    const bool       onGpu = a.device == "gpu";
    if (onGpu) {
        for (int i = 0; i < a.nDev; ++i)
            devs[i] = a.sameGpu ? 0 : i;
    }
    Neon::Backend bk(devs, onGpu ? Neon::Runtime::stream : Neon::Runtime::openmp);
    Lattice          lattice(bk);
    Neon::index_3d   dim(a.nx, a.ny, a.nz);

    Grid grid(
        bk, dim, [](const Neon::index_3d&) { return true; }, lattice.c_vect);

    PopulationField pop0 = grid.template newField<FP, Lattice::Q>("Population", Lattice::Q, FP(0.0));
    PopulationField pop1 = grid.template newField<FP, Lattice::Q>("Population", Lattice::Q, FP(0.0));
    CellType        defaultCelltype;
    auto            flag = grid.template newField<CellType, 1>("Material", 1, defaultCelltype);

    // omega exactly as Config.cpp:105-111 with N := nx (cubes in every golden case)
    const double nu = a.ulb * static_cast<double>(a.nx - 2) / a.Re;
    const double omegaD = 1. / (3. * nu + 0.5);
    const FP     omega = static_cast<FP>(omegaD);

    LbmIterationD3Q19<PopulationField, FP> iteration(Neon::set::StencilSemantic::standard,
                                                     a.occ ? Neon::skeleton::Occ::standard : Neon::skeleton::Occ::none,
                                                     Neon::set::TransferMode::get,
                                                     pop0, pop1, flag, omega);

    const auto&  t = lattice.t_vect;
    const auto&  c = lattice.c_vect;
    const double ulb = a.ulb;

    auto isEdge = [&](const Neon::index_3d& p) {
        return p.x == 0 || p.x == dim.x - 1 || p.y == 0 || p.y == dim.y - 1 || p.z == 0 || p.z == dim.z - 1;
    };
    auto initPop = [&](const Neon::index_3d& p, const int& k, FP& val) {
        val = t.at(k);
        if (isEdge(p)) {
            if (p.y == dim.y - 1) {
                val = -6. * t.at(k) * ulb * (c.at(k).v[0] * 1.0 + c.at(k).v[1] * 0.0 + c.at(k).v[2] * 0.0);
            } else {
                val = 0;
            }
        } else if (isSolidSphere(a, p.x, p.y, p.z)) {
            val = 0;
        }
    };
    iteration.getInput().forEachActiveCell(initPop);
    iteration.getOutput().forEachActiveCell(initPop);
    flag.forEachActiveCell([&](const Neon::index_3d& p, const int&, CellType& f) {
        f.classification = CellType::bulk;
        f.wallNghBitflag = 0;
        if (isEdge(p)) {
            f.classification = CellType::bounceBack;
            if (p.y == dim.y - 1) {
                f.classification = CellType::movingWall;
            }
        } else if (isSolidSphere(a, p.x, p.y, p.z)) {
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
    {
        auto container = LbmContainers<Lattice, PopulationField, FP>::computeWallNghMask(flag, flag);
        container.run(Neon::Backend::mainStreamIdx);
        bk.syncAll();
    }

    const int warm = a.warmup < 0 ? 0 : a.warmup;
    auto      start = std::chrono::high_resolution_clock::now();
    for (int it = 0; it < a.iters; ++it) {
        if (it == warm) {
            bk.syncAll();
            start = std::chrono::high_resolution_clock::now();
        }
        iteration.run();
    }
    bk.syncAll();
    auto stop = std::chrono::high_resolution_clock::now();
    if (a.warmup >= 0) {
        const double us = std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count();
        const double cells = double(a.nx) * a.ny * a.nz;
        const int    timed = a.iters - warm;
        std::printf(
            "{\"ref_bench\": true, \"grid\": \"%s\", \"fp\": \"%s\", \"nx\": %d, \"ny\": %d, \"nz\": %d, \"timed_iters\": %d, "
            "\"elapsed_us\": %.1f, \"mlups\": %.6f, \"ndev\": %d, \"device\": \"%s\"}\n",
            a.grid.c_str(), a.isDouble ? "double" : "float", a.nx, a.ny, a.nz, timed, us, cells * timed / us, a.nDev,
            a.device.c_str());
    }

    if (!a.dump.empty()) {
        auto& f = iteration.getInput();
        f.updateHostData(Neon::Backend::mainStreamIdx);
        flag.updateHostData(Neon::Backend::mainStreamIdx);
        bk.syncAll();
        const size_t          cells = size_t(a.nx) * a.ny * a.nz;
        std::vector<FP>       pops(cells * Lattice::Q);
        std::vector<uint32_t> mask(cells);
        std::vector<int32_t>  cls(cells);
        f.forEachActiveCell(
            [&](const Neon::index_3d& p, const int& k, FP& val) {
                pops[size_t(k) * cells + (size_t(p.z) * a.ny + p.y) * a.nx + p.x] = val;
            },
            Neon::computeMode_t::seq);
        flag.forEachActiveCell(
            [&](const Neon::index_3d& p, const int&, CellType& v) {
                const size_t o = (size_t(p.z) * a.ny + p.y) * a.nx + p.x;
                mask[o] = v.wallNghBitflag;
                cls[o] = static_cast<int32_t>(v.classification);
            },
            Neon::computeMode_t::seq);
        FILE* fp = std::fopen(a.dump.c_str(), "wb");
        if (!fp) {
            std::perror("dump");
            return 2;
        }
        int32_t hdr[8] = {0x4E4C424D, a.nx, a.ny, a.nz, Lattice::Q, int32_t(sizeof(FP)), a.iters, a.geom};
        std::fwrite(hdr, sizeof(hdr), 1, fp);
        std::fwrite(&omegaD, sizeof(double), 1, fp);
        std::fwrite(pops.data(), sizeof(FP), pops.size(), fp);
        std::fwrite(mask.data(), sizeof(uint32_t), mask.size(), fp);
        std::fwrite(cls.data(), sizeof(int32_t), cls.size(), fp);
        std::fclose(fp);
    }
    return 0;
}

int main(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i];
        auto        next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
        if (k == "--n") {
            a.nx = a.ny = a.nz = std::atoi(next());
        } else if (k == "--nx") {
            a.nx = std::atoi(next());
        } else if (k == "--ny") {
            a.ny = std::atoi(next());
        } else if (k == "--nz") {
            a.nz = std::atoi(next());
        } else if (k == "--iters") {
            a.iters = std::atoi(next());
        } else if (k == "--bench") {
            a.warmup = std::atoi(next());
        } else if (k == "--ndev") {
            a.nDev = std::atoi(next());
        } else if (k == "--geom") {
            a.geom = std::string(next()) == "sphere" ? 1 : 0;
        } else if (k == "--fp") {
            a.isDouble = std::string(next()) == "double";
        } else if (k == "--grid") {
            a.grid = next();
        } else if (k == "--dump") {
            a.dump = next();
        } else if (k == "--device") {
            a.device = next();
        } else if (k == "--occ") {
            a.occ = std::string(next()) == "standard";
        } else if (k == "--same-gpu") {
            a.sameGpu = true;
        } else {
            std::fprintf(stderr, "unknown arg %s\n", k.c_str());
            return 1;
        }
    }
    Neon::init();
    if (a.grid == "dGrid") {
        return a.isDouble ? runCase<Neon::dGrid, double>(a) : runCase<Neon::dGrid, float>(a);
    }
    if (a.grid == "bGrid") {
        return a.isDouble ? runCase<Neon::bGrid, double>(a) : runCase<Neon::bGrid, float>(a);
    }
    std::fprintf(stderr, "unknown grid %s\n", a.grid.c_str());
    return 1;
}
