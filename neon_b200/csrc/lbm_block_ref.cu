// lbm_block_ref.cu — NLBM_ARITH_REFERENCE instantiations of the block-sparse step kernel (-fmad=false: every operation
// rounds as the reference's CPU build does), plus the block-sparse set-up kernels, which must round the same way.
#include "lbm_block.cuh"
#include "lbm_host.h"

namespace nlbm {
cudaError_t launchBlockStepRef(StepKind kind, const BlockArgs& a, uint32_t nBlocks, cudaStream_t st, bool exact)
{
    switch (kind) {
        case kD3Q19_F32:
            if (exact)  // the same bits with a third of the float<->double conversions (lbm_collide_exact.cuh)
                return launchBlockStep<CollideD3Q19Exact<0>, float>(a, nBlocks, st);
            return launchBlockStep<CollideD3Q19Ref<float, float, 0>, float>(a, nBlocks, st);
        case kD3Q19_F64:
            return launchBlockStep<CollideD3Q19Ref<double, double, 0>, double>(a, nBlocks, st);
        case kD3Q19_F32C64:
            return launchBlockStep<CollideD3Q19Ref<float, double, 0>, float>(a, nBlocks, st);
        case kD3Q27_F32:
            return launchBlockStep<CollideD3Q27Ref<float, 0>, float>(a, nBlocks, st);
        case kD3Q27_F64:
            return launchBlockStep<CollideD3Q27Ref<double, 0>, double>(a, nBlocks, st);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ set-up (RunCavityTwoPop.cu:159-242 on a bGrid)
struct BlockGeom
{
    uint32_t*       flags;
    const uint32_t* info;
    uint32_t        nBlocksAlloc;
    int32_t         gnx, gny, gnz, geom;
    double          cx, cy, cz, R;
};

// one CTA of 512 threads per block (local AND ghost blocks, so that masks can look across partition faces)
__global__ void __launch_bounds__(512) k_block_classify(const BlockGeom g, const uint32_t* __restrict__ activeMask)
{
    const uint32_t blk = blockIdx.x;
    const int      t = threadIdx.x, lx = t & 7, ly = (t >> 3) & 7, lz = t >> 6;
    const uint32_t* inf = g.info + (int64_t)blk * 32;
    const int      x = (int)inf[27] + lx, y = (int)inf[28] + ly, z = (int)inf[29] + lz;
    uint32_t       c = NLBM_UNDEFINED;
    const bool     active = activeMask ? ((activeMask[(int64_t)blk * 16 + (t >> 5)] >> (t & 31)) & 1u) : true;
    if (active && x < g.gnx && y < g.gny && z < g.gnz) {
        c = NLBM_BULK;
        const bool   edge = x == 0 || x == g.gnx - 1 || y == 0 || y == g.gny - 1 || z == 0 || z == g.gnz - 1;
        const double dx = x - g.cx, dy = y - g.cy, dz = z - g.cz;
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
            if (y == 0 || y == g.gny - 1 || z == 0 || z == g.gnz - 1 || x == g.gnx - 1)
                c = NLBM_BOUNCE_BACK;
        }
    }
    g.flags[(int64_t)blk * kBlockCells + t] = c << NLBM_FLAG_CLASS_SHIFT;
}

// LbmContainers::computeWallNghMask (LbmTools.h:344-376) through the block connectivity; local blocks only
template <int Q>
__global__ void __launch_bounds__(512) k_block_wall_mask(const BlockGeom g, int32_t* __restrict__ bad)
{
    using L = Lattice<Q>;
    const uint32_t blk = blockIdx.x;
    const int      t = threadIdx.x, lx = t & 7, ly = (t >> 3) & 7, lz = t >> 6;
    const uint32_t* inf = g.info + (int64_t)blk * 32;
    const int64_t  o = (int64_t)blk * kBlockCells + t;
    const uint32_t cls = flagClass(g.flags[o]);
    uint32_t       m = 0;
    int            nbad = 0;
    if (cls == NLBM_BULK) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (q == L::REST)
                continue;
            const int xn = lx - L::c(q, 0), yn = ly - L::c(q, 1), zn = lz - L::c(q, 2);
            const int fx = (xn < 0) ? -1 : (xn >= kB ? 1 : 0), fy = (yn < 0) ? -1 : (yn >= kB ? 1 : 0), fz = (zn < 0) ? -1 : (zn >= kB ? 1 : 0);
            const uint32_t bn = (fx | fy | fz) ? inf[(fx + 1) + 3 * (fy + 1) + 9 * (fz + 1)] : blk;
            if (bn == kNoBlock) {
                ++nbad;  // the reference counts a missing neighbour as bulk (CellType.h:13-18) and then reads invalid data
                continue;
            }
            const uint32_t fn = g.flags[(int64_t)bn * kBlockCells + ((zn - fz * kB) * 64 + (yn - fy * kB) * 8 + (xn - fx * kB))];
            if (flagClass(fn) == NLBM_UNDEFINED) {
                ++nbad;  // a cell that is not active: same trap
                continue;
            }
            if (flagClass(fn) != NLBM_BULK)
                m |= 1u << q;
        }
    }
    uint32_t w = (cls << NLBM_FLAG_CLASS_SHIFT) | m;
    // a block whose 512 cells are all bulk without a wall neighbour is marked in the flag word of its first cell: the step kernel
    // then skips the block's flag words (any other writer of flag words — classify, an upload — rewrites that word without the bit)
    const int allPlain = __syncthreads_and(w == kPlainBulk);
    if (t == 0 && allPlain)
        w |= NLBM_FLAG_BLOCK_PLAIN;
    g.flags[o] = w;
    if (nbad && bad)
        atomicAdd(bad, nbad);
}

template <typename S, int Q>
__global__ void __launch_bounds__(512) k_block_init_pop(S* __restrict__ pop, const uint32_t* __restrict__ flags, int64_t popPitch, double ulb)
{
    using L = Lattice<Q>;
    const int64_t  o = (int64_t)blockIdx.x * kBlockCells + threadIdx.x;
    const uint32_t cls = flagClass(flags[o]);
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        S v = S(0);
        if (cls == NLBM_BULK) {
            v = (S)L::w(k);
        } else if (cls == NLBM_MOVING_WALL) {
            if constexpr (Q == 19) {  // RunCavityTwoPop.cu:177-184
                const double t = L::w(k);
                v = (S)(-6. * t * ulb * (L::c(k, 0) * 1. + L::c(k, 1) * 0. + L::c(k, 2) * 0.));
            } else {  // apps/lbmMultiRes/lidDrivenCavity.h:56-76 (same expressions as the dense k_init_pop)
                const double uw[3] = {ulb, 0., 0.};
                v = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    v += L::c(k, d) * uw[d];
                v *= -6. * L::w(k);
            }
        }
        pop[k * popPitch + o] = v;
    }
}

cudaError_t launchBlockClassify(const nlbm_block_desc& d, int geom, const double* sphere, const uint32_t* activeMask, cudaStream_t st)
{
    BlockGeom g;
    g.flags = d.flags;
    g.info = d.info;
    g.nBlocksAlloc = d.n_blocks_alloc;
    g.gnx = d.gnx;
    g.gny = d.gny;
    g.gnz = d.gnz;
    g.geom = geom;
    if (sphere) {
        g.cx = sphere[0];
        g.cy = sphere[1];
        g.cz = sphere[2];
        g.R = sphere[3];
    } else {
        const int mn = d.gnx < d.gny ? (d.gnx < d.gnz ? d.gnx : d.gnz) : (d.gny < d.gnz ? d.gny : d.gnz);
        g.cx = 0.45 * d.gnx;
        g.cy = 0.55 * d.gny;
        g.cz = 0.5 * d.gnz;
        g.R = mn / 5.0;
    }
    if (d.n_blocks_alloc == 0)
        return cudaSuccess;
    k_block_classify<<<d.n_blocks_alloc, 512, 0, st>>>(g, activeMask);
    return cudaGetLastError();
}

cudaError_t launchBlockWallMask(const nlbm_block_desc& d, int q, int32_t* bad, cudaStream_t st)
{
    BlockGeom g{};
    g.flags = d.flags;
    g.info = d.info;
    g.nBlocksAlloc = d.n_blocks_alloc;
    if (d.n_blocks == 0)
        return cudaSuccess;
    if (q == 19)
        k_block_wall_mask<19><<<d.n_blocks, 512, 0, st>>>(g, bad);
    else
        k_block_wall_mask<27><<<d.n_blocks, 512, 0, st>>>(g, bad);
    return cudaGetLastError();
}

template <typename S>
cudaError_t launchBlockInitPop(const nlbm_block_desc& d, int q, double ulb, cudaStream_t st)
{
    if (d.n_blocks_alloc == 0)
        return cudaSuccess;
    const int64_t pitch = (int64_t)d.n_blocks_alloc * kBlockCells;
    if (q == 19)
        k_block_init_pop<S, 19><<<d.n_blocks_alloc, 512, 0, st>>>((S*)d.pop_out, d.flags, pitch, ulb);
    else
        k_block_init_pop<S, 27><<<d.n_blocks_alloc, 512, 0, st>>>((S*)d.pop_out, d.flags, pitch, ulb);
    return cudaGetLastError();
}
template cudaError_t launchBlockInitPop<float>(const nlbm_block_desc&, int, double, cudaStream_t);
template cudaError_t launchBlockInitPop<double>(const nlbm_block_desc&, int, double, cudaStream_t);

}  // namespace nlbm
