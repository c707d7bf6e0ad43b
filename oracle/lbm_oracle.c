/* TEST INFRASTRUCTURE — CPU oracle for the Neon LBM hot path.
 *
 * A plain-C restatement of the reference algorithm (Autodesk/Neon v0.3.3),
 * written from the cited lines, NOT shipped and NOT on any product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Parity status: PINNED for D3Q19 (dGrid, fp32 and fp64,
 * cavity and cavity+sphere) against the unmodified reference built by
 * oracle/Makefile.ref — see tests/golden/ and oracle/make_golden.py.  D3Q27 is
 * "parity unpinned": upstream never runs it on a uniform grid and
 * apps/lbmMultiRes does not build offline (glm/libigl), so it is pinned only
 * by this restatement of apps/lbmMultiRes/{lattice,stream,collide,util}.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* threads used by the step loops: 1 by default (the scalar port that bench.py may time as "cores": 1);
 * tests raise it so that hundreds of iterations finish in seconds — results are bit-identical. */
static int g_olbm_threads = 1;
void       olbm_set_threads(int n) { g_olbm_threads = n < 1 ? 1 : n; }
int        olbm_get_threads(void) { return g_olbm_threads; }

enum { OLBM_BOUNCE = 0, OLBM_MOVING = 1, OLBM_BULK = 2 }; /* src/CellType.h:5-11 */

/* benchmarks/lbm-lid-driven-cavity-flow/src/D3Q19.h:23-44 (velocities), :112-132 (weights) */
static const int OLBM_C19[19][3] = {
    {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {0, -1, -1}, {0, -1, 1},
    {0, 0, 0},
    {1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 0},   {1, -1, 0}, {1, 0, 1},   {1, 0, -1}, {0, 1, 1},   {0, 1, -1}};
static const double OLBM_W19[19] = {1. / 18., 1. / 18., 1. / 18., 1. / 36., 1. / 36., 1. / 36., 1. / 36.,
                                    1. / 36., 1. / 36., 1. / 3.,  1. / 18., 1. / 18., 1. / 18., 1. / 36.,
                                    1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36.};

/* apps/lbmMultiRes/lattice.h:15-77 */
static const int OLBM_C27[27][3] = {
    {0, 0, 0},   {0, 0, -1},   {0, 0, 1},   {0, -1, 0}, {0, -1, -1}, {0, -1, 1}, {0, 1, 0},  {0, 1, -1},  {0, 1, 1},
    {-1, 0, 0},  {-1, 0, -1},  {-1, 0, 1},  {-1, -1, 0}, {-1, -1, -1}, {-1, -1, 1}, {-1, 1, 0}, {-1, 1, -1}, {-1, 1, 1},
    {1, 0, 0},   {1, 0, -1},   {1, 0, 1},   {1, -1, 0}, {1, -1, -1}, {1, -1, 1}, {1, 1, 0},  {1, 1, -1},  {1, 1, 1}};
static const int OLBM_OPP27[27] = {0,  2,  1,  6,  8,  7,  3,  5,  4,  18, 20, 19, 24, 26,
                                   25, 21, 23, 22, 9,  11, 10, 15, 17, 16, 12, 14, 13};
static const double OLBM_W27[27] = {
    8.0 / 27.0, 2.0 / 27.0,  2.0 / 27.0,  2.0 / 27.0, 1.0 / 54.0,  1.0 / 54.0,  2.0 / 27.0, 1.0 / 54.0,  1.0 / 54.0,
    2.0 / 27.0, 1.0 / 54.0,  1.0 / 54.0,  1.0 / 54.0, 1.0 / 216.0, 1.0 / 216.0, 1.0 / 54.0, 1.0 / 216.0, 1.0 / 216.0,
    2.0 / 27.0, 1.0 / 54.0,  1.0 / 54.0,  1.0 / 54.0, 1.0 / 216.0, 1.0 / 216.0, 1.0 / 54.0, 1.0 / 216.0, 1.0 / 216.0};

void olbm_tables(int q_lat, int* c_out, int* opp_out, double* w_out)
{
    for (int q = 0; q < q_lat; ++q) {
        for (int d = 0; d < 3; ++d)
            c_out[3 * q + d] = q_lat == 19 ? OLBM_C19[q][d] : OLBM_C27[q][d];
        opp_out[q] = q_lat == 19 ? (q == 9 ? 9 : (q < 9 ? q + 10 : q - 10)) : OLBM_OPP27[q];
        w_out[q] = q_lat == 19 ? OLBM_W19[q] : OLBM_W27[q];
    }
}

/* Cell classes.  geom 0: lid-driven cavity, RunCavityTwoPop.cu:208-224 — the shell of the box is
 * bounceBack, except y = ny-1 which is movingWall (takes precedence).
 * geom 1: the same box with a solid (bounceBack) sphere, the obstacle definition of
 * oracle/ref_driver.cu (centre (0.45nx, 0.55ny, 0.5nz), R = min(n)/5).
 * geom 2: flow over sphere as defined for the uniform build in SURVEY.md §8d from
 * apps/lbmMultiRes/flowOverShape.h:64-100,165-175: x=0 plane is the inlet (handled exactly like the
 * moving wall), y/z extreme planes and x=nx-1 are bounceBack, sphere centre/radius given.   */
void olbm_classify(int geom, int nx, int ny, int nz, const double* sphere /*cx,cy,cz,R or NULL*/, int32_t* cls)
{
    double cx = 0.45 * nx, cy = 0.55 * ny, cz = 0.5 * nz;
    int    m = nx < ny ? nx : ny;
    m = m < nz ? m : nz;
    double R = m / 5.0;
    if (sphere) {
        cx = sphere[0];
        cy = sphere[1];
        cz = sphere[2];
        R = sphere[3];
    }
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t o = ((size_t)z * ny + y) * nx + x;
                int32_t      c = OLBM_BULK;
                const int    edge = x == 0 || x == nx - 1 || y == 0 || y == ny - 1 || z == 0 || z == nz - 1;
                if (geom == 0 || geom == 1) {
                    if (edge) {
                        c = OLBM_BOUNCE;
                        if (y == ny - 1)
                            c = OLBM_MOVING;
                    } else if (geom == 1) {
                        const double dx = x - cx, dy = y - cy, dz = z - cz;
                        if (dx * dx + dy * dy + dz * dz < R * R)
                            c = OLBM_BOUNCE;
                    }
                } else {
                    const double dx = x - cx, dy = y - cy, dz = z - cz;
                    if (x == 0)
                        c = OLBM_MOVING;
                    if (dx * dx + dy * dy + dz * dz < R * R)
                        c = OLBM_BOUNCE;
                    if (y == 0 || y == ny - 1 || z == 0 || z == nz - 1 || x == nx - 1)
                        c = OLBM_BOUNCE;
                }
                cls[o] = c;
            }
}

/* Wall-neighbour mask: LbmTools.h:327-376.  Only bulk cells get a mask; bit k set <=> the cell at
 * x - c_k is not bulk.  A neighbour outside the domain counts as bulk (CellType.h:13-18: the int
 * ctor ignores its argument).  Returns the number of bulk cells that have a
 * neighbour outside the domain (the reference would then read invalid data; must be 0).   */
long olbm_wall_mask(int q_lat, int nx, int ny, int nz, const int32_t* cls, uint32_t* mask)
{
    long bad = 0;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t o = ((size_t)z * ny + y) * nx + x;
                uint32_t     m = 0;
                if (cls[o] == OLBM_BULK) {
                    for (int q = 0; q < q_lat; ++q) {
                        const int* c = q_lat == 19 ? OLBM_C19[q] : OLBM_C27[q];
                        if (c[0] == 0 && c[1] == 0 && c[2] == 0)
                            continue;
                        const int xn = x - c[0], yn = y - c[1], zn = z - c[2];
                        if (xn < 0 || xn >= nx || yn < 0 || yn >= ny || zn < 0 || zn >= nz) {
                            ++bad;
                            continue;
                        }
                        if (cls[((size_t)zn * ny + yn) * nx + xn] != OLBM_BULK)
                            m |= 1u << q;
                    }
                }
                mask[o] = m;
            }
    return bad;
}

#define STORE float
#define COMPUTE float
#define SFX(n) n##_f32
#include "lbm_oracle_impl.h"
#undef STORE
#undef COMPUTE
#undef SFX

#define STORE double
#define COMPUTE double
#define SFX(n) n##_f64
#include "lbm_oracle_impl.h"
#undef STORE
#undef COMPUTE
#undef SFX

/* store float / compute double: the "f/d" column of the reference sweep
 * (lbm-lid-driven-cavity-flow.py:1-10) */
#define STORE float
#define COMPUTE double
#define SFX(n) n##_f32c64
#include "lbm_oracle_impl.h"
#undef STORE
#undef COMPUTE
#undef SFX
