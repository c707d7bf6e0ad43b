// TEST INFRASTRUCTURE — not part of the product path.
//
// Pins D3Q27 to code compiled from the UNMODIFIED reference (Autodesk/Neon v0.3.3, apps/lbmMultiRes), which is where
// D3Q27 lives upstream.  This file is our own driver; everything numerical it runs is the
// reference's, compiled from where it lies under /root/reference with -DKBC (which selects the 27-velocity tables):
//   latticeVelocity / latticeOppositeID / latticeWeights   apps/lbmMultiRes/lattice.h:15-77
//   stream<T, Q>        (pull, half-way bounce-back rule)   apps/lbmMultiRes/stream.h:5-49
//   collideBGK<T, Q>    (generic BGK)                       apps/lbmMultiRes/collide.h:286-354
//   velocity<T, Q>, computeOmega, getDir                    apps/lbmMultiRes/util.h:13-62
// on a ONE-level Neon::domain::mGrid (the uniform case of the multi-resolution grid, CPU backend).  glm is not available
// offline; oracle/ref_stubs/glm stands in for it (util.h needs it only for a geometry helper that is never called here).
// postProcess.h / lbmMultiRes.h (polyscope, Eigen, libigl) are not included, so the cavity set-up of
// lidDrivenCavity.h:30-76 is restated below with the reference's own tables.
//
// Order of the two containers per iteration: stream, then collide — the fused pull kernel of the dGrid benchmark
// (pullStream -> collide, LbmTools.h:301-322) that BASELINE.json configs[4] combines D3Q27 with.  (The multi-resolution
// app itself runs collide first; on one level that is the same recurrence shifted by half an iteration.)
//
//   ref_lbm27 --tables                         prints the lattice tables as JSON
//   ref_lbm27 --n NX NY NZ --iters K --fp double|float --geom 0|1 --dump FILE
// Dump format: as oracle/ref_driver.cu (magic, nx, ny, nz, q, fp_bytes, iters, geom, omega, populations [q][z][y][x],
// wall masks [z][y][x] computed with the stream.h rule, classes [z][y][x]).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "Neon/Neon.h"
#include "Neon/domain/mGrid.h"

#include "lattice.h"
#include "util.h"
#include "collide.h"
#include "stream.h"

constexpr int Q = 27;

struct Args
{
    int         nx = 12, ny = 12, nz = 12, iters = 5, geom = 0;
    bool        isDouble = true, tables = false;
    std::string dump;
    double      Re = 100., ulb = 0.04, omega = -1.;
};

static bool isSolidSphere(const Args& a, int x, int y, int z)
{
    if (a.geom != 1)
        return false;
    // the obstacle of oracle/ref_driver.cu: centre (0.45nx, 0.55ny, 0.5nz), R = min(n)/5
    const double cx = 0.45 * a.nx, cy = 0.55 * a.ny, cz = 0.5 * a.nz;
    int          m = a.nx < a.ny ? a.nx : a.ny;
    m = m < a.nz ? m : a.nz;
    const double R = m / 5.0;
    const double dx = x - cx, dy = y - cy, dz = z - cz;
    return dx * dx + dy * dy + dz * dz < R * R;
}

// The device tables are `__device__ static constexpr`: readable in constant expressions on the host
// (--expt-relaxed-constexpr), which is how this prints the very arrays the kernels index.
template <int I>
static void printRow(FILE* f)
{
    constexpr int    cx = latticeVelocity[I][0], cy = latticeVelocity[I][1], cz = latticeVelocity[I][2], op = latticeOppositeID[I];
    constexpr double w = latticeWeights[I];
    fprintf(f, "%s{\"c\": [%d, %d, %d], \"opp\": %d, \"w\": %.17g}", I ? ", " : "", cx, cy, cz, op, w);
    if constexpr (I + 1 < Q)
        printRow<I + 1>(f);
}

template <typename T>
static int run(const Args& a)
{
    const std::vector<int> devs(1, 0);
    Neon::Backend        backend(devs, Neon::Runtime::openmp);
    const Neon::index_3d dim(a.nx, a.ny, a.nz);
    const Neon::mGridDescriptor<1> descriptor(1);
    std::vector<std::function<bool(const Neon::index_3d&)>> active(1);
    active[0] = [](const Neon::index_3d&) { return true; };
    Neon::domain::mGrid grid(backend, dim, active, Neon::domain::Stencil::s19_t(false), descriptor);

    auto fin = grid.newField<T>("fin", Q, 0);
    auto fout = grid.newField<T>("fout", Q, 0);
    auto cellType = grid.newField<CellType>("CellType", 1, CellType::bulk);

    // cavity set-up: lidDrivenCavity.h:30-76 (classes :42-53, populations :56-76), plus the optional solid sphere
    const Neon::double_3d ulid(a.ulb, 0., 0.);
    cellType.forEachActiveCell(0, [&](const Neon::index_3d& idx, const int&, CellType& t) {
        t = CellType::bulk;
        if (idx.x == 0 || idx.x == a.nx - 1 || idx.y == 0 || idx.y == a.ny - 1 || idx.z == 0 || idx.z == a.nz - 1) {
            t = CellType::bounceBack;
            if (idx.y == a.ny - 1)
                t = CellType::movingWall;
        } else if (isSolidSphere(a, idx.x, idx.y, idx.z)) {
            t = CellType::bounceBack;
        }
    }, false, Neon::computeMode_t::computeMode_e::seq);
    struct Row { int c[3]; double w; };
    std::vector<Row> rows;
    {   // host copies of the reference tables, taken through constant expressions
        auto fill = [&](auto self, auto I) -> void {
            constexpr int i = decltype(I)::value;
            constexpr int cx = latticeVelocity[i][0], cy = latticeVelocity[i][1], cz = latticeVelocity[i][2];
            constexpr double w = latticeWeights[i];
            rows.push_back(Row{{cx, cy, cz}, w});
            if constexpr (i + 1 < Q)
                self(self, std::integral_constant<int, i + 1>{});
        };
        fill(fill, std::integral_constant<int, 0>{});
    }
    auto initPop = [&](const Neon::index_3d& idx, const int& q, T& v) {
        const CellType t = cellType(idx, 0, 0);
        T pop_init_val = rows[q].w;
        if (t == CellType::bounceBack)
            pop_init_val = 0;
        if (t == CellType::movingWall) {
            pop_init_val = 0;
            for (int d = 0; d < 3; ++d)
                pop_init_val += rows[q].c[d] * ulid.v[d];
            pop_init_val *= -6. * rows[q].w;
        }
        v = pop_init_val;
    };
    fin.forEachActiveCell(0, initPop, false, Neon::computeMode_t::computeMode_e::seq);
    fout.forEachActiveCell(0, initPop, false, Neon::computeMode_t::computeMode_e::seq);
    cellType.updateDeviceData();
    fin.updateDeviceData();
    fout.updateDeviceData();

    // omega: lidDrivenCavity.h:243-246 (clength = N of the coarsest level; one level: N) unless given
    const T clength = T(a.nx);
    const T visclb = T(a.ulb) * clength / static_cast<T>(a.Re);
    const T omega = a.omega > 0 ? T(a.omega) : T(1.0 / (3. * visclb + 0.5));

    auto S = stream<T, Q>(grid, 0, cellType, fout, fin);                 // fin(x, q) <- pull from fout
    auto C = collideBGK<T, Q>(grid, omega, 0, 1, cellType, fin, fout);  // fout <- BGK(fin)
    for (int it = 0; it < a.iters; ++it) {
        S.run(0);
        C.run(0);
    }
    backend.syncAll();
    fout.updateHostData();

    if (!a.dump.empty()) {
        const size_t          cells = (size_t)a.nx * a.ny * a.nz;
        std::vector<T>        pop(cells * Q);
        std::vector<int32_t>  cls(cells);
        std::vector<uint32_t> mask(cells, 0);
        for (int z = 0; z < a.nz; ++z)
            for (int y = 0; y < a.ny; ++y)
                for (int x = 0; x < a.nx; ++x) {
                    const size_t         o = ((size_t)z * a.ny + y) * a.nx + x;
                    const Neon::index_3d idx(x, y, z);
                    cls[o] = (int32_t)cellType(idx, 0, 0);
                    for (int q = 0; q < Q; ++q)
                        pop[(size_t)q * cells + o] = fout(idx, q, 0);
                }
        // wall bits as stream.h:28-43 decides them: bit q <=> the cell at x - c_q is not bulk
        for (int z = 0; z < a.nz; ++z)
            for (int y = 0; y < a.ny; ++y)
                for (int x = 0; x < a.nx; ++x) {
                    const size_t o = ((size_t)z * a.ny + y) * a.nx + x;
                    if (cls[o] != CellType::bulk)
                        continue;
                    for (int q = 1; q < Q; ++q) {
                        const int xn = x - rows[q].c[0], yn = y - rows[q].c[1], zn = z - rows[q].c[2];
                        if (xn < 0 || yn < 0 || zn < 0 || xn >= a.nx || yn >= a.ny || zn >= a.nz)
                            continue;
                        if (cls[((size_t)zn * a.ny + yn) * a.nx + xn] != CellType::bulk)
                            mask[o] |= 1u << q;
                    }
                }
        FILE* f = fopen(a.dump.c_str(), "wb");
        if (!f)
            return 2;
        const int32_t hdr[8] = {0x4E4C424D, a.nx, a.ny, a.nz, Q, (int32_t)sizeof(T), a.iters, a.geom};
        const double  om = (double)omega;
        fwrite(hdr, sizeof hdr, 1, f);
        fwrite(&om, sizeof om, 1, f);
        fwrite(pop.data(), sizeof(T), pop.size(), f);
        fwrite(mask.data(), sizeof(uint32_t), mask.size(), f);
        fwrite(cls.data(), sizeof(int32_t), cls.size(), f);
        fclose(f);
    }
    printf("{\"ref27\": true, \"n\": [%d, %d, %d], \"iters\": %d, \"fp\": \"%s\", \"omega\": %.17g}\n", a.nx, a.ny, a.nz, a.iters,
           sizeof(T) == 8 ? "double" : "float", (double)omega);
    return 0;
}

int main(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string k = argv[i];
        if (k == "--tables")
            a.tables = true;
        else if (k == "--n" && i + 3 < argc) {
            a.nx = atoi(argv[++i]);
            a.ny = atoi(argv[++i]);
            a.nz = atoi(argv[++i]);
        } else if (k == "--iters" && i + 1 < argc)
            a.iters = atoi(argv[++i]);
        else if (k == "--geom" && i + 1 < argc)
            a.geom = atoi(argv[++i]);
        else if (k == "--omega" && i + 1 < argc)
            a.omega = atof(argv[++i]);
        else if (k == "--fp" && i + 1 < argc)
            a.isDouble = std::string(argv[++i]) == "double";
        else if (k == "--dump" && i + 1 < argc)
            a.dump = argv[++i];
        else {
            fprintf(stderr, "unknown argument %s\n", k.c_str());
            return 1;
        }
    }
    if (a.tables) {
        printf("{\"q\": %d, \"rows\": [", Q);
        printRow<0>(stdout);
        printf("]}\n");
        return 0;
    }
    Neon::init();
    return a.isDouble ? run<double>(a) : run<float>(a);
}
