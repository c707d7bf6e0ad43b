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
#include <utility>
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
template <typename Grid>
void testAxpy(const Neon::Backend& bk, int n, const char* gridName)
{
    using Field = typename Grid::template Field<double, 3>;
    /* block-sparse grids: only a ball of cells is active, so the active mask decides which threads run the lambda */
    const bool           sparse = std::is_same_v<Grid, Neon::bGrid>;
    const Neon::index_3d dim(n, n + 3, std::max(n + 1, 16 * bk.getDeviceCount()));
    auto                 active = [=](const Neon::index_3d& p) {
        const double dx = p.x - 0.5 * dim.x, dy = p.y - 0.5 * dim.y, dz = p.z - 0.5 * dim.z;
        return !sparse || dx * dx + dy * dy + 0.25 * dz * dz < 0.2 * dim.x * dim.x;
    };
    Grid grid(bk, dim, active, Neon::domain::Stencil());
    Field       x = grid.template newField<double, 3>("x", 3, 0.0), y = grid.template newField<double, 3>("y", 3, 0.0);
    x.forEachActiveCell([](const Neon::index_3d& p, const int& c, double& v) { v = 0.25 * p.x - 0.5 * p.y + p.z + c; });
    y.forEachActiveCell([](const Neon::index_3d& p, const int& c, double& v) { v = 1.0 + p.x * p.y - c * p.z; });
    x.updateDeviceData();
    y.updateDeviceData();
    const double a = 1.5;
    auto         axpy = grid.newContainer("axpy", [&](Neon::set::Loader& L) {
        const auto& xp = L.load(const_cast<const Field&>(x));
        auto&       yp = L.load(y);
        return [=] NEON_CUDA_HOST_DEVICE(const typename Grid::Idx& i) mutable {
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
    size_t visited = 0;
    y.forEachActiveCell([&](const Neon::index_3d&, const int& c, double&) { visited += c == 0; }, Neon::computeMode_t::seq);
    report((std::string("MAP axpy on ") + gridName + " (" + std::to_string(visited) + " active cells)").c_str(),
           err == 0.0 && visited == grid.getNumActiveCells(), err);
    if (axpy.getTokens().size() != 2 || axpy.getTokens()[0].access != Neon::set::Access::read ||
        axpy.getTokens()[1].access != Neon::set::Access::write) {
        report("MAP tokens (const field = read, non-const = write)", false, 0);
    }
}

// ------------------------------------------------------------------------------------------------------ 2. STENCIL
template <typename Grid>
void testDiffusion(const Neon::Backend& bk, int n, Neon::skeleton::Occ occ, Neon::set::TransferMode mode, const char* gridName)
{
    using Field = typename Grid::template Field<double, 1>;
    const std::vector<Neon::index_3d> star = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    const Neon::index_3d              dim(n + 2, n, std::max(n + 5, 16 * bk.getDeviceCount() + 3));
    Grid                              grid(bk, dim, [](const Neon::index_3d&) { return true; }, star);
    Field                             u[2] = {grid.template newField<double, 1>("u0", 1, 0.0), grid.template newField<double, 1>("u1", 1, 0.0)};
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
            return [=] NEON_CUDA_HOST_DEVICE(const typename Grid::Idx& i) mutable {
                const double c = a(i, 0);
                /* missing neighbours (outside the box) take the cell's own value: zero-flux walls */
                const double s = a.template getNghData<-1, 0, 0>(i, 0, c) + a.template getNghData<1, 0, 0>(i, 0, c) +
                                 a.template getNghData<0, -1, 0>(i, 0, c) + a.template getNghData<0, 1, 0>(i, 0, c) +
                                 a.template getNghData<0, 0, -1>(i, 0, c) + a.template getNghData<0, 0, 1>(i, 0, c);
                b(i, 0) = c + k * (s - 6.0 * c);
            };
        });
        sk[t].sequence(std::vector<Neon::set::Container>{step}, "diffusion", Neon::skeleton::Options(occ, mode));
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
    const std::string name = std::string("STENCIL diffusion on ") + gridName + ", OCC " + Neon::skeleton::OccUtils::toString(occ) + ", " +
                             Neon::set::TransferModeUtils::toString(mode);
    report(name.c_str(), err <= 1e-13 * scale, err / scale);
}

// ------------------------------------------------------------------------------------- 2b. map -> stencil -> map sequence
// One Skeleton holding three containers (the shape Occ::extended / Occ::twoWayExtended transform, multiGpuGraph.cpp:145-301):
// b = 2a + 1 (map), c = z-stencil of b (stencil), a = 0.25 c - b (map; feeds the next run).  Compared with the host.
void testSequence(const Neon::Backend& bk, int n, Neon::skeleton::Occ occ)
{
    using Field = Neon::dGrid::Field<double, 1>;
    const std::vector<Neon::index_3d> zstar = {{0, 0, -1}, {0, 0, 1}};
    const Neon::index_3d              dim(n, n / 2 + 1, std::max(n + 3, 6 * bk.getDeviceCount()));
    Neon::dGrid                       grid(bk, dim, [](const Neon::index_3d&) { return true; }, zstar);
    Field                             a = grid.newField<double, 1>("a", 1, 0.0), b = grid.newField<double, 1>("b", 1, 0.0), c = grid.newField<double, 1>("c", 1, 0.0);
    auto                              init = [](const Neon::index_3d& p, const int&, double& v) { v = 0.01 * p.x - 0.02 * p.y + 0.001 * p.z * p.z; };
    a.forEachActiveCell(init);
    a.updateDeviceData();
    auto m1 = grid.newContainer("M1", [&](Neon::set::Loader& L) {
        const auto& ap = L.load(const_cast<const Field&>(a));
        auto&       bp = L.load(b);
        return [=] NEON_CUDA_HOST_DEVICE(const Neon::dGrid::Idx& i) mutable { bp(i, 0) = 2.0 * ap(i, 0) + 1.0; };
    });
    auto st = grid.newContainer("S", [&](Neon::set::Loader& L) {
        const auto& bp = L.load(const_cast<const Field&>(b), Neon::Pattern::STENCIL);
        auto&       cp = L.load(c);
        return [=] NEON_CUDA_HOST_DEVICE(const Neon::dGrid::Idx& i) mutable {
            cp(i, 0) = bp.template getNghData<0, 0, -1>(i, 0, 0.0) + 2.0 * bp(i, 0) + bp.template getNghData<0, 0, 1>(i, 0, 0.0);
        };
    });
    auto m2 = grid.newContainer("M2", [&](Neon::set::Loader& L) {
        const auto& cp = L.load(const_cast<const Field&>(c));
        const auto& bp = L.load(const_cast<const Field&>(b));
        auto&       ap = L.load(a);
        return [=] NEON_CUDA_HOST_DEVICE(const Neon::dGrid::Idx& i) mutable { ap(i, 0) = 0.25 * cp(i, 0) - bp(i, 0); };
    });
    Neon::skeleton::Skeleton sk(bk);
    sk.sequence({m1, st, m2}, "map-stencil-map", Neon::skeleton::Options(occ, Neon::set::TransferMode::get));
    const int runs = 5;
    for (int r = 0; r < runs; ++r) {
        sk.run();
    }
    a.updateHostData();
    bk.syncAll();
    const size_t        cells = dim.rMul<size_t>();
    std::vector<double> A(cells), B(cells), C(cells);
    auto                at = [&](std::vector<double>& f, int x, int y, int z) -> double& { return f[(size_t(z) * dim.y + y) * dim.x + x]; };
    for (int z = 0; z < dim.z; ++z)
        for (int y = 0; y < dim.y; ++y)
            for (int x = 0; x < dim.x; ++x) {
                int k = 0;
                init(Neon::index_3d(x, y, z), k, at(A, x, y, z));
            }
    for (int r = 0; r < runs; ++r) {
        for (size_t i = 0; i < cells; ++i)
            B[i] = 2.0 * A[i] + 1.0;
        for (int z = 0; z < dim.z; ++z)
            for (int y = 0; y < dim.y; ++y)
                for (int x = 0; x < dim.x; ++x)
                    at(C, x, y, z) = (z > 0 ? at(B, x, y, z - 1) : 0.0) + 2.0 * at(B, x, y, z) + (z + 1 < dim.z ? at(B, x, y, z + 1) : 0.0);
        for (size_t i = 0; i < cells; ++i)
            A[i] = 0.25 * C[i] - B[i];
    }
    double err = 0, scale = 0;
    a.forEachActiveCell(
        [&](const Neon::index_3d& p, const int&, double& v) {
            err = std::max(err, std::fabs(v - at(A, p.x, p.y, p.z)));
            scale = std::max(scale, std::fabs(at(A, p.x, p.y, p.z)));
        },
        Neon::computeMode_t::seq);
    const std::string name = std::string("SEQUENCE map -> stencil -> map in one Skeleton, OCC ") + Neon::skeleton::OccUtils::toString(occ);
    report(name.c_str(), err <= 1e-12 * scale, err / scale);
}

// ---------------------------------------------------------------------------------------------------------- 3. LBM
// A D3Q19 pull-stream + BGK step written as a USER lambda, the way an application author writes it against Neon's API
// (benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h:99-282 does the same by macro expansion): one statement per
// population with COMPILE-TIME neighbour offsets (getNghData<dx,dy,dz>), so the launcher's address arithmetic folds to
// immediates on the padded pitch.  It is the workload that measures the generic launch path against the reference's.
struct Lattice19
{
    /* D3Q19.h:23-44 numbering, opposite of k is k +- 10, rest population 9 */
    NEON_CUDA_HOST_DEVICE static constexpr int c(int k, int d)
    {
        constexpr int t[19][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {0, -1, -1}, {0, -1, 1}, {0, 0, 0},
                                  {1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 0},   {1, -1, 0}, {1, 0, 1},   {1, 0, -1}, {0, 1, 1},   {0, 1, -1}};
        return t[k][d];
    }
    NEON_CUDA_HOST_DEVICE static constexpr int   opp(int k) { return k == 9 ? 9 : (k < 9 ? k + 10 : k - 10); }
    NEON_CUDA_HOST_DEVICE static constexpr float w(int k) { return k == 9 ? 1.f / 3.f : ((k % 10) < 3 ? 1.f / 18.f : 1.f / 36.f); }
};

template <int K, typename Part>
NEON_CUDA_HOST_DEVICE inline float pullOne(const Part& in, const Neon::dGrid::Idx& i, uint32_t word)
{
    using L = Lattice19;
    constexpr int bx = -L::c(K, 0), by = -L::c(K, 1), bz = -L::c(K, 2);
    if (K != 9 && (word & (1u << K))) { /* the cell at x - c_k is a wall: half-way bounce-back (+ lid momentum) */
        return in(i, L::opp(K)) + in.template getNghData<bx, by, bz>(i, L::opp(K)).mData;
    }
    return in.template getNghData<bx, by, bz>(i, K).mData;
}
template <typename Part, int... Ks>
NEON_CUDA_HOST_DEVICE inline void pullAll(std::integer_sequence<int, Ks...>, const Part& in, const Neon::dGrid::Idx& i, uint32_t word, float (&f)[19])
{
    ((f[Ks] = pullOne<Ks>(in, i, word)), ...);
}
template <typename Part, int... Ks>
NEON_CUDA_HOST_DEVICE inline void collideAll(std::integer_sequence<int, Ks...>, Part& out, const Neon::dGrid::Idx& i, const float (&f)[19], float rho,
                                             float ux, float uy, float uz, float usqr, float omega)
{
    using L = Lattice19;
    ((out(i, Ks) = (1.f - omega) * f[Ks] +
                   omega * (rho * L::w(Ks) *
                            (1.f + 3.f * (L::c(Ks, 0) * ux + L::c(Ks, 1) * uy + L::c(Ks, 2) * uz) +
                             4.5f * (L::c(Ks, 0) * ux + L::c(Ks, 1) * uy + L::c(Ks, 2) * uz) * (L::c(Ks, 0) * ux + L::c(Ks, 1) * uy + L::c(Ks, 2) * uz) - usqr))),
     ...);
}

template <typename Pop, typename Flag>
Neon::set::Container userLbmStep(const Neon::dGrid& grid, Neon::set::StencilSemantic semantic, const Pop& fIn, const Flag& flag, float omega,
                                 Pop& fOut)
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
            pullAll(std::make_integer_sequence<int, 19>{}, in, i, word, f);
            float rho = 0, ux = 0, uy = 0, uz = 0;
#pragma unroll
            for (int k = 0; k < 19; ++k) {
                rho += f[k];
                ux += f[k] * Lattice19::c(k, 0);
                uy += f[k] * Lattice19::c(k, 1);
                uz += f[k] * Lattice19::c(k, 2);
            }
            ux /= rho;
            uy /= rho;
            uz /= rho;
            const float usqr = 1.5f * (ux * ux + uy * uy + uz * uz);
            collideAll(std::make_integer_sequence<int, 19>{}, out, i, f, rho, ux, uy, uz, usqr, omega);
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
    for (int k = 0; k < 19; ++k) { /* the compile-time table of the user lambda is the benchmark's lattice */
        for (int d = 0; d < 3; ++d) {
            if (Lattice19::c(k, d) != lattice.c_vect[k].v[d]) {
                report("lattice table of the user lambda", false, 0);
            }
        }
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
    user[0].sequence({userLbmStep(grid, sem, const_cast<const Pop&>(a0), flag, omega, a1)}, "user0", Neon::skeleton::Options(occ, Neon::set::TransferMode::get));
    user[1].sequence({userLbmStep(grid, sem, const_cast<const Pop&>(a1), flag, omega, a0)}, "user1", Neon::skeleton::Options(occ, Neon::set::TransferMode::get));
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
            testAxpy<Neon::dGrid>(bk, n, "dGrid");
            testAxpy<Neon::bGrid>(bk, n, "bGrid");
            for (auto occ : {Neon::skeleton::Occ::none, Neon::skeleton::Occ::standard, Neon::skeleton::Occ::extended, Neon::skeleton::Occ::twoWayExtended}) {
                for (auto mode : {Neon::set::TransferMode::get, Neon::set::TransferMode::put}) {
                    testDiffusion<Neon::dGrid>(bk, n, occ, mode, "dGrid");
                    testDiffusion<Neon::bGrid>(bk, n, occ, mode, "bGrid");
                }
                testSequence(bk, n, occ);
                testLbm(bk, n, occ, 0);
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return failures ? 1 : 0;
}
