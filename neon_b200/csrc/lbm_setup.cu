// lbm_setup.cu — problem set-up and output kernels of the dense path (compiled with -fmad=false so that the
// geometry predicates and initial values round exactly as the reference's host code does).
//
//   classify   : RunCavityTwoPop.cu:208-224 (cavity), apps/lbmMultiRes/flowOverShape.h:64-100,165-175 (sphere)
//   wall mask  : LbmContainers::computeWallNghMask, LbmTools.h:344-376   (bit-exact)
//   init pops  : RunCavityTwoPop.cu:168-206, apps/lbmMultiRes/lidDrivenCavity.h:56-76
//   rho / u    : LbmContainers::computeRhoAndU, LbmTools.h:384-437
#include "lbm_common.cuh"
#include "lbm_host.h"

namespace nlbm {

struct GeomArgs
{
    uint32_t* flags;
    int32_t   nx, ny, nzm, pitch_y;
    int64_t   pitch_z;
    int32_t   z_origin, z_halo, gnx, gny, gnz;
    int32_t   geom;
    double    cx, cy, cz, R;
};

__global__ void k_classify(const GeomArgs g)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zm = blockIdx.z;
    if (x >= g.pitch_y)
        return;
    const int gz = g.z_origin + zm - g.z_halo;
    uint32_t  c = NLBM_UNDEFINED;
    if (x < g.nx && gz >= 0 && gz < g.gnz) {
        c = NLBM_BULK;
        const bool   edge = x == 0 || x == g.gnx - 1 || y == 0 || y == g.gny - 1 || gz == 0 || gz == g.gnz - 1;
        const double dx = x - g.cx, dy = y - g.cy, dz = gz - g.cz;
        const bool   inSphere = dx * dx + dy * dy + dz * dz < g.R * g.R;
        if (g.geom == 0 || g.geom == 1) {
            if (edge) {
                c = NLBM_BOUNCE_BACK;
                if (y == g.gny - 1)
                    c = NLBM_MOVING_WALL;
            } else if (g.geom == 1 && inSphere) {
                c = NLBM_BOUNCE_BACK;
            }
        } else {
            if (x == 0)
                c = NLBM_MOVING_WALL;
            if (inSphere)
                c = NLBM_BOUNCE_BACK;
            if (y == 0 || y == g.gny - 1 || gz == 0 || gz == g.gnz - 1 || x == g.gnx - 1)
                c = NLBM_BOUNCE_BACK;
        }
    }
    g.flags[(int64_t)zm * g.pitch_z + (int64_t)y * g.pitch_y + x] = c << NLBM_FLAG_CLASS_SHIFT;
}

template <int Q>
__global__ void k_wall_mask(const GeomArgs g, int32_t* __restrict__ bad)
{
    using L = Lattice<Q>;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zm = blockIdx.z + g.z_halo;  // local planes only
    if (x >= g.nx)
        return;
    const int      gz = g.z_origin + zm - g.z_halo;
    const int64_t  o = (int64_t)zm * g.pitch_z + (int64_t)y * g.pitch_y + x;
    const uint32_t cls = flagClass(g.flags[o]);
    uint32_t       m = 0;
    int            nbad = 0;
    if (cls == NLBM_BULK) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (q == L::REST)
                continue;
            const int xn = x - L::c(q, 0), yn = y - L::c(q, 1), zn = gz - L::c(q, 2);
            if (xn < 0 || xn >= g.gnx || yn < 0 || yn >= g.gny || zn < 0 || zn >= g.gnz) {
                ++nbad;  // the reference treats a missing neighbour as bulk (CellType.h:13-18) and then reads invalid data
                continue;
            }
            const int zmn = zm - L::c(q, 2);
            if (zmn < 0 || zmn >= g.nzm) {
                ++nbad;  // neighbour in another partition but no ghost plane to look at
                continue;
            }
            const uint32_t fn = g.flags[(int64_t)zmn * g.pitch_z + (int64_t)yn * g.pitch_y + xn];
            if (flagClass(fn) != NLBM_BULK)
                m |= 1u << q;
        }
    }
    // In place (RunCavityTwoPop.cu:239 runs it on (flag, flag)): other threads read THIS word's class bits while the mask
    // bits change.  Only the mask bits are touched, by atomic read-modify-writes, so the class bits (and any bit of the
    // word the caller uses for itself) are never rewritten and a concurrent reader always sees them whole.
    atomicAnd(&g.flags[o], ~kMaskBits);
    if (m)
        atomicOr(&g.flags[o], m);
    if (nbad && bad)
        atomicAdd(bad, nbad);
}

// one warp per (row, summary word): 32 chunks of 32 cells; also writes the cell map (one byte per 4 cells, lbm_common.cuh)
__global__ void k_summary(const uint32_t* __restrict__ flags, uint2* __restrict__ summary, uint8_t* __restrict__ cellMap, int nx, int ny,
                          int nzm, int pitch_y, int64_t pitch_z, int wpr)
{
    const int     lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t rows = (int64_t)ny * nzm;
    if (wid >= rows * wpr)
        return;
    const int64_t row = wid / wpr;
    const int     word = (int)(wid % wpr);
    const int     zm = (int)(row / ny), y = (int)(row % ny);
    const uint32_t* r = flags + (int64_t)zm * pitch_z + (int64_t)y * pitch_y;
    uint32_t        spec = 0, bulk = 0;
    for (int c = 0; c < 32; ++c) {
        const int x = (word * 32 + c) * kChunk + lane;
        bool      isSpec = false, isBulk = false;
        if (x < pitch_y) {
            const uint32_t f = x < nx ? r[x] : (uint32_t)NLBM_UNDEFINED << NLBM_FLAG_CLASS_SHIFT;
            isBulk = flagIsBulk(f);
            isSpec = f != kPlainBulk;
        }
        const uint32_t ms = __ballot_sync(0xffffffffu, isSpec), mb = __ballot_sync(0xffffffffu, isBulk);
        if (ms)
            spec |= 1u << c;
        if (mb)
            bulk |= 1u << c;
        const int xg = (word * 32 + c) * kChunk;  // first cell of the chunk: lanes 0..7 write the bytes of its 8 groups of 4 cells
        if (lane < 8 && xg + 4 * lane < pitch_y)
            cellMap[row * (pitch_y >> 2) + (xg >> 2) + lane] = (uint8_t)(((mb >> (4 * lane)) & 0xFu) | (((ms >> (4 * lane)) & 0xFu) << 4));
    }
    if (lane == 0)
        summary[row * wpr + word] = make_uint2(spec, bulk);
}

template <typename S, int Q>
__global__ void k_init_pop(S* __restrict__ pop, const uint32_t* __restrict__ flags, int nx, int pitch_y, int64_t pitch_z,
                           int64_t pitch_q, double ulb)
{
    using L = Lattice<Q>;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= pitch_y)
        return;
    const int64_t  o = (int64_t)blockIdx.z * pitch_z + (int64_t)blockIdx.y * pitch_y + x;
    const uint32_t cls = x < nx ? flagClass(flags[o]) : (uint32_t)NLBM_UNDEFINED;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        S val = 0;
        if (cls == NLBM_BULK) {
            val = (S)L::w(q);
        } else if (cls == NLBM_MOVING_WALL) {
            if constexpr (Q == 19) {
                const double t = L::w(q);
                val = (S)(-6. * t * ulb * (L::c(q, 0) * 1. + L::c(q, 1) * 0. + L::c(q, 2) * 0.));
            } else {
                const double uw[3] = {ulb, 0., 0.};
                val = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    val += L::c(q, d) * uw[d];
                val *= -6. * L::w(q);
            }
        }
        pop[q * pitch_q + o] = val;
    }
}

template <typename S, typename C>
__global__ void k_rho_u(const S* __restrict__ in, const uint32_t* __restrict__ flags, S* __restrict__ rho_out,
                        S* __restrict__ u_out, int nx, int ny, int nzm, int pitch_y, int64_t pitch_z, int64_t pitch_q,
                        int z_halo)
{
    using L = Lattice<19>;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zm = blockIdx.z + z_halo;
    if (x >= nx)
        return;
    const int64_t  o = (int64_t)zm * pitch_z + (int64_t)y * pitch_y + x;
    const uint32_t f = flags[o];
    C              rho = 0, u[3] = {0, 0, 0};
    if (flagIsBulk(f)) {
        S p[19];
#pragma unroll
        for (int q = 0; q < 19; ++q) {
            const int64_t dn = L::c(q, 2) * pitch_z + (int64_t)L::c(q, 1) * pitch_y + L::c(q, 0);
            if (q != 9 && (f & (1u << q))) {
                const S* po = in + L::opp(q) * pitch_q + o;
                p[q] = po[0] + po[-dn];
            } else {
                p[q] = in[q * pitch_q + o - dn];
            }
        }
#define P(i) ((C)p[i])
        const C X_M1 = P(0) + P(3) + P(4) + P(5) + P(6);
        const C X_P1 = P(10) + P(13) + P(14) + P(15) + P(16);
        const C X_0 = P(9) + P(1) + P(2) + P(7) + P(8) + P(11) + P(12) + P(17) + P(18);
        const C Y_M1 = P(1) + P(3) + P(7) + P(8) + P(14);
        const C Y_P1 = P(4) + P(11) + P(13) + P(17) + P(18);
        const C Z_M1 = P(2) + P(5) + P(7) + P(16) + P(18);
        const C Z_P1 = P(6) + P(8) + P(12) + P(15) + P(17);
#undef P
        rho = X_M1 + X_P1 + X_0;
        u[0] = (X_P1 - X_M1) / rho;
        u[1] = (Y_P1 - Y_M1) / rho;
        u[2] = (Z_P1 - Z_M1) / rho;
    } else if (flagClass(f) == NLBM_MOVING_WALL) {
        rho = 1.0;
#pragma unroll
        for (int d = 0; d < 3; ++d)
            u[d] = (C)in[d * pitch_q + o] / (C)(6. * 1. / 18.);
    }
    rho_out[o] = (S)rho;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        u_out[d * pitch_q + o] = (S)u[d];
}

// ------------------------------------------------------------------ host side
static GeomArgs geomArgs(const nlbm_dense_desc& d, int geom, const double* sphere)
{
    GeomArgs g;
    g.flags = d.flags;
    g.nx = d.nx;
    g.ny = d.ny;
    g.nzm = d.nz_local + 2 * d.z_halo;
    g.pitch_y = (int32_t)d.pitch_y;
    g.pitch_z = d.pitch_z;
    g.z_origin = d.z_origin;
    g.z_halo = d.z_halo;
    g.gnx = d.gnx;
    g.gny = d.gny;
    g.gnz = d.gnz;
    g.geom = geom;
    int m = d.gnx < d.gny ? d.gnx : d.gny;
    m = m < d.gnz ? m : d.gnz;
    g.cx = 0.45 * d.gnx;
    g.cy = 0.55 * d.gny;
    g.cz = 0.5 * d.gnz;
    g.R = m / 5.0;
    if (sphere) {
        g.cx = sphere[0];
        g.cy = sphere[1];
        g.cz = sphere[2];
        g.R = sphere[3];
    }
    return g;
}

// x-face cache of a field: cache[side][q][zm][y] = field[q][zm][y][side ? nx-1 : 0]   (include/neon_lbm.h)
template <typename W>
__global__ void k_wall_cache_build(const W* __restrict__ field, W* __restrict__ cache, int nx, int ny, int nzm, int q, int64_t pitch_y,
                                   int64_t pitch_z, int64_t pitch_q)
{
    const int64_t rows = (int64_t)q * nzm * ny;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * rows)
        return;
    const int     side = t >= rows;
    const int64_t r = t - side * rows;
    const int     y = (int)(r % ny);
    const int     zm = (int)((r / ny) % nzm);
    const int     c = (int)(r / ((int64_t)ny * nzm));
    cache[t] = field[c * pitch_q + zm * pitch_z + y * pitch_y + (side ? nx - 1 : 0)];
}
cudaError_t launchWallCacheBuild(const nlbm_dense_desc& d, int q, int elemBytes, cudaStream_t st)
{
    const int     nzm = d.nz_local + 2 * d.z_halo;
    const int64_t n = (int64_t)2 * q * nzm * d.ny;
    const int     threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (elemBytes == 4)
        k_wall_cache_build<uint32_t><<<blocks, threads, 0, st>>>((const uint32_t*)d.pop_out, (uint32_t*)d.wall_cache, d.nx, d.ny, nzm, q,
                                                                 d.pitch_y, d.pitch_z, d.pitch_q);
    else
        k_wall_cache_build<uint64_t><<<blocks, threads, 0, st>>>((const uint64_t*)d.pop_out, (uint64_t*)d.wall_cache, d.nx, d.ny, nzm, q,
                                                                 d.pitch_y, d.pitch_z, d.pitch_q);
    return cudaGetLastError();
}

cudaError_t launchSummary(const nlbm_dense_desc& d, cudaStream_t st)
{
    const int     nzm = d.nz_local + 2 * d.z_halo;
    const int     wpr = (int)summaryWordsPerRow(d.pitch_y);
    const int64_t warps = (int64_t)d.ny * nzm * wpr;
    const int     threads = 256;
    const int64_t blocks = (warps * 32 + threads - 1) / threads;
    k_summary<<<(unsigned)blocks, threads, 0, st>>>(d.flags, const_cast<uint2*>(summaryPtr(d)), const_cast<uint8_t*>(cellMapPtr(d)), d.nx,
                                                    d.ny, nzm, (int)d.pitch_y, d.pitch_z, wpr);
    return cudaGetLastError();
}

// flag words from one byte per cell (FlagField::setClasses with a host mirror of classes): class bits set, wall bits
// cleared; padding cells and memory planes the caller has no classes for become `undefined`
__global__ void k_flags_from_classes(uint32_t* __restrict__ flags, const uint8_t* __restrict__ cls, int nx, int ny, int pitch_y,
                                     int64_t pitch_z, int zmFirst, int nPlanes)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zm = blockIdx.z;
    if (x >= pitch_y)
        return;
    uint32_t  c = NLBM_UNDEFINED;
    const int p = zm - zmFirst;
    if (x < nx && p >= 0 && p < nPlanes)
        c = cls[((int64_t)p * ny + y) * nx + x] & 3u;
    flags[(int64_t)zm * pitch_z + (int64_t)y * pitch_y + x] = c << NLBM_FLAG_CLASS_SHIFT;
}

cudaError_t launchFlagsFromClasses(const nlbm_dense_desc& d, const uint8_t* cls, int zmFirst, int nPlanes, cudaStream_t st)
{
    const int nzm = d.nz_local + 2 * d.z_halo;
    dim3      block(128), grid(((int)d.pitch_y + 127) / 128, d.ny, nzm);
    k_flags_from_classes<<<grid, block, 0, st>>>(const_cast<uint32_t*>(d.flags), cls, d.nx, d.ny, (int)d.pitch_y, d.pitch_z, zmFirst, nPlanes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    return launchSummary(d, st);
}

cudaError_t launchClassify(const nlbm_dense_desc& d, int geom, const double* sphere, cudaStream_t st)
{
    const GeomArgs g = geomArgs(d, geom, sphere);
    dim3           block(128), grid((g.pitch_y + 127) / 128, d.ny, g.nzm);
    k_classify<<<grid, block, 0, st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    return launchSummary(d, st);
}

cudaError_t launchWallMask(const nlbm_dense_desc& d, int q, int32_t* d_bad, cudaStream_t st)
{
    const GeomArgs g = geomArgs(d, 0, nullptr);
    dim3           block(128), grid((d.nx + 127) / 128, d.ny, d.nz_local);
    if (q == 19)
        k_wall_mask<19><<<grid, block, 0, st>>>(g, d_bad);
    else
        k_wall_mask<27><<<grid, block, 0, st>>>(g, d_bad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    return launchSummary(d, st);
}

template <typename S>
cudaError_t launchInitPop(const nlbm_dense_desc& d, int q, double ulb, cudaStream_t st)
{
    const int nzm = d.nz_local + 2 * d.z_halo;
    dim3      block(128), grid(((int)d.pitch_y + 127) / 128, d.ny, nzm);
    if (q == 19)
        k_init_pop<S, 19><<<grid, block, 0, st>>>((S*)d.pop_out, d.flags, d.nx, (int)d.pitch_y, d.pitch_z, d.pitch_q, ulb);
    else
        k_init_pop<S, 27><<<grid, block, 0, st>>>((S*)d.pop_out, d.flags, d.nx, (int)d.pitch_y, d.pitch_z, d.pitch_q, ulb);
    return cudaGetLastError();
}
template cudaError_t launchInitPop<float>(const nlbm_dense_desc&, int, double, cudaStream_t);
template cudaError_t launchInitPop<double>(const nlbm_dense_desc&, int, double, cudaStream_t);

template <typename S, typename C>
cudaError_t launchRhoU(const nlbm_dense_desc& d, void* rho, void* u, cudaStream_t st)
{
    const int nzm = d.nz_local + 2 * d.z_halo;
    dim3      block(128), grid((d.nx + 127) / 128, d.ny, d.nz_local);
    k_rho_u<S, C><<<grid, block, 0, st>>>((const S*)d.pop_in, d.flags, (S*)rho, (S*)u, d.nx, d.ny, nzm, (int)d.pitch_y,
                                          d.pitch_z, d.pitch_q, d.z_halo);
    return cudaGetLastError();
}
template cudaError_t launchRhoU<float, float>(const nlbm_dense_desc&, void*, void*, cudaStream_t);
template cudaError_t launchRhoU<double, double>(const nlbm_dense_desc&, void*, void*, cudaStream_t);

}  // namespace nlbm
