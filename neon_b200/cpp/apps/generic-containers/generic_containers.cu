// generic-containers — user-written per-cell device lambdas through Grid::newContainer on this library's layout
// (SURVEY.md §8f.4), checked against host loops and against the native fused LBM kernels.
//
//   generic-containers [--deviceIds 0 0 1 ...] [--n 48] [--bench N]
// 1. MAP      axpy over a 3-component double field                         == host loop, bit for bit
// 2. STENCIL  explicit diffusion step with getNghData<dx,dy,dz>(idx, c, alt), Skeleton + OCC, "grid" halo semantic
//                                                                          == host loop within 1e-13
// 3. LBM      D3Q19 pull + BGK written as a user lambda (runtime-offset getNghData, flag words), Skeleton + OCC with the
//             lattice halo semantic                                        == native nlbm_d3q19_f32_dense_step within 1e-5
// --bench N: times the user-lambda LBM against the native kernel on an N^3 cavity (what the hand-written kernel buys).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "Neon/Neon.h"
#include "Neon/domain/GenericContainer.h"
#include "Neon/domain/dGrid.h"
#include "Neon/lbm/Lbm.h"
#include "Neon/skeleton/Skeleton.h"

namespace generic_test {

int failures = 0;
void report(const char* what, bool ok, double err)
{
    std::printf("%s %s (max err %.3e)\n", ok ? "PASS" : "FAIL", what, err);
    failures += ok ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------------------- 1. MAP
void testAxpy(const Neon::Backend& bk, int n)
{
    using Field = Neon::dGrid::Field<double, 3>;
    Neon::dGrid grid(bk, {n, n + 3, n + 1}, [](const Neon::index_3d&) { return true; }, Neon::domain::Stencil());
    Field       x = grid.newField<double, 3>("x", 3, 0.0), y = grid.newField<double, 3>("y", 3, 0.0);
    x.forEachActiveCell([](const Neon::index_3d& p, const int& c, double& v) { v = 0.25 * p.x - 0.5 * p.y + p.z + c; });
    y.forEachActiveCell([](const Neon::index_3d& p, const int& c, double& v) { v = 1.0 + p.x * p.y - c * p.z; });
    x.updateDeviceData();
    y.updateDeviceData();
    const double a = 1.5;
    auto         axpy = grid.newContainer("axpy", [&](Neon::set::Loader& L) {
        const auto& xp = L.load(const_cast<const Field&>(x));
        auto&       yp = L.load(y);
        return [=] NEON_CUDA_HOST_DEVICE(const Neon::dGrid::Idx& i) mutable {
            for (int c = 0; c < yp.cardinality(); ++c) {
                yp(i, c) = yp(i, c) + a * xp(i, c);
            }
        };
    });
    axpy.run(Neon::Backend::mainStreamIdx);
    y.updateHostData();
    bk.syncAll();
    double err = 0;
    y.forEachActiveCell(
        [&](const Neon::index_3d& p, const int& c, double& v) {
            const double want = (1.0 + p.x * p.y - c * p.z) + a * (0.25 * p.x - 0.5 * p.y + p.z + c);
            err = std::max(err, std::fabs(v - want));
        },
        Neon::computeMode_t::seq);
    report("MAP axpy", err == 0.0, err);
    if (axpy.getTokens().size() != 2 || axpy.getTokens()[0].access != Neon::set::Access::read ||
        axpy.getTokens()[1].access != Neon::set::Access::write) {
        report("MAP tokens (const field = read, non-const = write)", false, 0);
    }
}

// ------------------------------------------------------------------------------------------------------ 2. STENCIL
void testDiffusion(const Neon::Backend& bk, int n, Neon::skeleton::Occ occ, Neon::set::TransferMode mode)
{
    using Field = Neon::dGrid::Field<double, 1>;
    const std::vector<Neon::index_3d> star = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    const Neon::index_3d              dim(n + 2, n, n + 5);
    Neon::dGrid                       grid(bk, dim, [](const Neon::index_3d&) { return true; }, star);
    Field                             u[2] = {grid.newField<double, 1>("u0", 1, 0.0), grid.newField<double, 1>("u1", 1, 0.0)};
    auto                              init = [](const Neon::index_3d& p, const int&, double& v) { v = std::sin(0.3 * p.x) + 0.1 * p.y * p.z; };
    u[0].forEachActiveCell(init);
    u[0].updateDeviceData();
    const double               k = 0.1;
    Neon::skeleton::Skeleton   sk[2] = {Neon::skeleton::Skeleton(bk), Neon::skeleton::Skeleton(bk)};
    for (int t = 0; t < 2; ++t) {
        const Field& in = u[t];
        Field&       out = u[1 - t];
        auto         step = grid.newContainer("diffusion", [&](Neon::set::Loader& L) {
            const auto& a = L.load(in, Neon::Pattern::STENCIL);
            auto&       b = L.load(out);
            return [=] NEON_CUDA_HOST_DEVICE(const Neon::dGrid::Idx& i) mutable {
                const double c = a(i, 0);
                /* missing neighbours (outside the box) take the cell's own value: zero-flux walls */
                const double s = a.template getNghData<-1, 0, 0>(i, 0, c) + a.template getNghData<1, 0, 0>(i, 0, c) +
                                 a.template getNghData<0, -1, 0>(i, 0, c) + a.template getNghData<0, 1, 0>(i, 0, c) +
                                 a.template getNghData<0, 0, -1>(i, 0, c) + a.template getNghData<0, 0, 1>(i, 0, c);
                b(i, 0) = c + k * (s - 6.0 * c);
            };
        });
        sk[t].sequence({step}, "diffusion", Neon::skeleton::Options(occ, mode));
    }
    const int iters = 7;
    for (int it = 0; it < iters; ++it) {
        sk[it & 1].run();
    }
    Field& res = u[iters & 1];
    res.updateHostData();
    bk.syncAll();
    /* host reference */
    const size_t        cells = dim.rMul<size_t>();
    std::vector<double> a(cells), b(cells);
    auto                at = [&](std::vector<double>& f, int x, int y, int z) -> double& { return f[(size_t(z) * dim.y + y) * dim.x + x]; };
    for (int z = 0; z < dim.z; ++z)
        for (int y = 0; y < dim.y; ++y)
            for (int x = 0; x < dim.x; ++x) {
                int c = 0;
                init(Neon::index_3d(x, y, z), c, at(a, x, y, z));
            }
    for (int it = 0; it < iters; ++it) {
        for (int z = 0; z < dim.z; ++z)
            for (int y = 0; y < dim.y; ++y)
                for (int x = 0; x < dim.x; ++x) {
                    const double c = at(a, x, y, z);
                    auto         ngh = [&](int dx, int dy, int dz) {
                        const int X = x + dx, Y = y + dy, Z = z + dz;
                        return (X < 0 || Y < 0 || Z < 0 || X >= dim.x || Y >= dim.y || Z >= dim.z) ? c : at(a, X, Y, Z);
                    };
                    const double s = ngh(-1, 0, 0) + ngh(1, 0, 0) + ngh(0, -1, 0) + ngh(0, 1, 0) + ngh(0, 0, -1) + ngh(0, 0, 1);
                    at(b, x, y, z) = c + k * (s - 6.0 * c);
                }
        a.swap(b);
    }
    double err = 0, scale = 0;
    res.forEachActiveCell(
        [&](const Neon::index_3d& p, const int&, double& v) {
            err = std::max(err, std::fabs(v - at(a, p.x, p.y, p.z)));
            scale = std::max(scale, std::fabs(at(a, p.x, p.y, p.z)));
        },
        Neon::computeMode_t::seq);
    const std::string name = std::string("STENCIL diffusion, OCC ") + Neon::skeleton::OccUtils::toString(occ) + ", " +
                             Neon::set::TransferModeUtils::toString(mode);
    report(name.c_str(), err <= 1e-13 * scale, err / scale);
}

// ---------------------------------------------------------------------------------------------------------- 3. LBM
struct Lattice19
{
    int8_t c[19][3];
    float  w[19];
};

template <typename Pop, typename Flag>
Neon::set::Container userLbmStep(const Neon::dGrid& grid, Neon::set::StencilSemantic semantic, const Pop& fIn, const Flag& flag, float omega,
                                 Pop& fOut, const Lattice19& lat)
{
    return grid.newContainer("userLambdaLBM", [&](Neon::set::Loader& L) {
        const auto& in = L.load(fIn, Neon::Pattern::STENCIL, semantic);
        const auto& fl = L.load(flag);
        auto&       out = L.load(fOut);
        return [=] NEON_CUDA_HOST_DEVICE(const Neon::dGrid::Idx& i) mutable {
            const uint32_t word = fl(i, 0);
            if (NLBM_FLAG_CLASS(word) != NLBM_BULK) {
                return; /* non-bulk cells are never written (LbmTools.h:304) */
            }
            float f[19];
            for (int k = 0; k < 19; ++k) {
                const int            opp = k == 9 ? 9 : (k < 9 ? k + 10 : k - 10);
                const Neon::index_3d back(-lat.c[k][0], -lat.c[k][1], -lat.c[k][2]);
                if (word & (1u << k)) { /* the cell at x - c_k is a wall: half-way bounce-back (+ lid momentum) */
                    f[k] = in(i, opp) + in.getNghData(i, back, opp).mData;
                } else {
                    f[k] = in.getNghData(i, back, k).mData;
                }
            }
            float rho = 0, ux = 0, uy = 0, uz = 0;
            for (int k = 0; k < 19; ++k) {
                rho += f[k];
                ux += f[k] * lat.c[k][0];
                uy += f[k] * lat.c[k][1];
                uz += f[k] * lat.c[k][2];
            }
            ux /= rho;
            uy /= rho;
            uz /= rho;
            const float usqr = 1.5f * (ux * ux + uy * uy + uz * uz);
            for (int k = 0; k < 19; ++k) {
                const float cu = 3.f * (lat.c[k][0] * ux + lat.c[k][1] * uy + lat.c[k][2] * uz);
                const float eq = rho * lat.w[k] * (1.f + cu + 0.5f * cu * cu - usqr);
                out(i, k) = (1.f - omega) * f[k] + omega * eq;
            }
        };
    });
}

void testLbm(const Neon::Backend& bk, int n, Neon::skeleton::Occ occ, int benchIters)
{
    using Lattice = D3Q19Template<float, float>;
    using Pop = Neon::dGrid::Field<float, 19>;
    Lattice              lattice(bk);
    const Neon::index_3d dim(n, n, n);
    Neon::dGrid          grid(bk, dim, [](const Neon::index_3d&) { return true; }, lattice.c_vect);
    Pop                  a0 = grid.newField<float, 19>("a0", 19, 0.f), a1 = grid.newField<float, 19>("a1", 19, 0.f);
    Pop                  b0 = grid.newField<float, 19>("b0", 19, 0.f), b1 = grid.newField<float, 19>("b1", 19, 0.f);
    auto                 flag = grid.newField<CellType, 1>("flag", 1, CellType());
    Lattice19            lat{};
    for (int k = 0; k < 19; ++k) {
        for (int d = 0; d < 3; ++d) {
            lat.c[k][d] = int8_t(lattice.c_vect[k].v[d]);
        }
        lat.w[k] = float(lattice.t_vect[k]);
    }
    const double nu = 0.04 * double(n - 2) / 100.0;
    const float  omega = float(1. / (3. * nu + 0.5));
    /* device-side set-up of the cavity (with the obstacle when the box is small enough to be a parity case) */
    for (int d = 0; d < bk.getDeviceCount(); ++d) {
        bk.setDevice(d);
        nlbm_dense_desc desc = a0.getPartition(d).desc;
        desc.flags = flag.getPartition(d).mem();
        cudaStream_t st = bk.stream(d, 0);
        Neon::detail::check(nlbm_dense_classify(&desc, benchIters ? 0 : 1, nullptr, st), "classify");
        Neon::detail::check(nlbm_dense_wall_mask(&desc, 19, nullptr, st), "wall mask");
        for (Pop* f : {&a0, &a1, &b0, &b1}) {
            desc.pop_out = f->getPartition(d).mem();
            Neon::detail::check(nlbm_dense_init_pop_f32(&desc, 19, 0.04, st), "init");
        }
    }
    for (Pop* f : {&a0, &a1, &b0, &b1}) {
        f->commitWalls(); /* a0/a1 lose theirs again as soon as the user lambda writes them; b0/b1 keep it for the native kernel */
    }
    bk.syncAll();
    const auto               sem = Neon::set::StencilSemantic::streaming;
    Neon::skeleton::Skeleton user[2] = {Neon::skeleton::Skeleton(bk), Neon::skeleton::Skeleton(bk)};
    user[0].sequence({userLbmStep(grid, sem, const_cast<const Pop&>(a0), flag, omega, a1, lat)}, "user0", Neon::skeleton::Options(occ, Neon::set::TransferMode::get));
    user[1].sequence({userLbmStep(grid, sem, const_cast<const Pop&>(a1), flag, omega, a0, lat)}, "user1", Neon::skeleton::Options(occ, Neon::set::TransferMode::get));
    LbmIterationD3Q19<Pop, float> native(sem, occ, Neon::set::TransferMode::get, b0, b1, flag, omega);
    const int iters = benchIters ? benchIters : 20;
    auto      timeIt = [&](auto&& body) {
        bk.syncAll();
        const auto t0 = std::chrono::high_resolution_clock::now();
        body();
        bk.syncAll();
        return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    };
    if (benchIters) { /* warm-up */
        for (int it = 0; it < 4; ++it) {
            user[it & 1].run();
            native.run();
        }
    }
    const double tUser = timeIt([&] {
        for (int it = 0; it < iters; ++it) {
            user[it & 1].run();
        }
    });
    const double tNative = timeIt([&] {
        for (int it = 0; it < iters; ++it) {
            native.run();
        }
    });
    Pop& ru = (iters & 1) ? a1 : a0;
    Pop& rn = native.getInput();
    ru.updateHostData();
    rn.updateHostData();
    bk.syncAll();
    double err = 0, scale = 0;
    const float* pu = ru.hostData();
    const float* pn = rn.hostData();
    for (size_t o = 0; o < dim.rMul<size_t>() * 19; ++o) {
        err = std::max(err, double(std::fabs(pu[o] - pn[o])));
        scale = std::max(scale, double(std::fabs(pn[o])));
    }
    const std::string name = std::string("LBM user lambda vs native kernel, OCC ") + Neon::skeleton::OccUtils::toString(occ);
    report(name.c_str(), err <= 1e-5 * scale, err / scale);
    if (benchIters) {
        const double cells = double(dim.rMul<size_t>()) * iters;
        std::printf("{\"generic_bench\": true, \"n\": %d, \"devices\": %d, \"iters\": %d, \"user_lambda_mlups\": %.1f, \"native_mlups\": %.1f}\n", n,
                    bk.getDeviceCount(), iters, cells / tUser * 1e-6, cells / tNative * 1e-6);
    }
}

}  // namespace generic_test
using namespace generic_test;

int main(int argc, char** argv)
{
    std::vector<int> devs;
    int              n = 40, bench = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string k = argv[i];
        if (k == "--deviceIds") {
            while (i + 1 < argc && std::isdigit(argv[i + 1][0])) {
                devs.push_back(std::atoi(argv[++i]));
            }
        } else if (k == "--n" && i + 1 < argc) {
            n = std::atoi(argv[++i]);
        } else if (k == "--bench" && i + 1 < argc) {
            bench = std::atoi(argv[++i]);
        } else {
            std::fprintf(stderr, "usage: %s [--deviceIds id...] [--n N] [--bench N]\n", argv[0]);
            return 2;
        }
    }
    if (devs.empty()) {
        devs.push_back(0);
    }
    try {
        Neon::init();
        Neon::Backend bk(devs, Neon::Runtime::stream);
        if (bench) {
            testLbm(bk, bench, Neon::skeleton::Occ::standard, 50);
        } else {
            testAxpy(bk, n);
            for (auto occ : {Neon::skeleton::Occ::none, Neon::skeleton::Occ::standard}) {
                for (auto mode : {Neon::set::TransferMode::get, Neon::set::TransferMode::put}) {
                    testDiffusion(bk, n, occ, mode);
                }
                testLbm(bk, n, occ, 0);
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return failures ? 1 : 0;
}
