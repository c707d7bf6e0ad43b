// lbm_block.cuh — fused pull-stream + BGK collide over BLOCK-SPARSE fields (Neon's bGrid: 8 x 8 x 8-cell blocks).
//
// Replaces the generic lambda kernel over a bSpan (libNeonDomain/include/Neon/domain/details/bGrid/bSpan_imp.h:7-21,
// libNeonSet/include/Neon/set/LambdaExecutor.h:100-160) carrying LbmContainers::iteration (LbmTools.h:285-325), whose
// every neighbour access resolves a block through the 27-entry connectivity table and tests the neighbour's active bit
// (bPartition_imp.h:97-124, 194-198, 218-309, 340-358; StaticBlock.h:47-103).
//
// Layout (B200: SoA per population, every block's population tile is one contiguous, 2 KB-aligned run):
//   pop[q][blk][z][y][x]      element offset (q * n_blocks_alloc + blk) * 512 + z*64 + y*8 + x
//   flags[blk][z][y][x]       the dense path's flag word; cells that are not active carry class UNDEFINED
//   info[blk][32]             27 neighbour block ids, index (dx+1) + 3(dy+1) + 9(dz+1) as bPartition_imp.h:194-198
//                             (NLBM_NO_BLOCK if absent), then the block origin x, y, z and two spare words: ONE 128-byte line
// (reference: [blk][q][z][y][x] with 32-bit offsets, connectivity / origin / active-mask in three separate arrays.)
//
// Kernel: one CTA per block, one thread per VEC consecutive cells of a row (fp32: VEC = 4, 128 threads).  As in the
// dense kernel every load is a predicated volatile PTX load issued before anything is consumed.  The block's info line
// gates the neighbour addresses; it is prefetched into L2 by the CTA that ran ~one chip-load of blocks earlier, so that it
// costs an L2 hit, not a DRAM round trip, in front of the population loads.  (A first version issued the rows inside the
// block before the info line arrived and the others after: ptxas serialises the two predicated loads of a row on their
// shared destination registers, ncu r01i.)  The x shift inside a row is a warp shuffle.  Fix-ups, collision and the
// single store per population are the dense kernel's (finishCells semantics) with block-aware addressing.
#pragma once
#include "lbm_step.cuh"

namespace nlbm {

constexpr int      kB = 8;                  // block edge (Neon::bGrid = StaticBlock<8,8,8>, domain/bGrid.h:5)
constexpr int      kBlockCells = kB * kB * kB;
constexpr uint32_t kNoBlock = NLBM_NO_BLOCK;
constexpr uint32_t kPrefetchAhead = 148 * 6;  // ~ the CTAs resident on the chip: whose info line to pull into L2

struct BlockArgs
{
    const void*     in;
    void*           out;
    const uint32_t* flags;
    const uint32_t* info;
    uint32_t        firstBlock;  // the view's first block
    uint32_t        nBlocks;     // blocks of this launch
    int64_t         popPitch;    // elements between populations = n_blocks_alloc * 512
    int32_t         eagerFlags;  // NLBM_OPT_FLAG_WORDS: fetch every flag word with the populations, ignore the block marker
    double          omega;
};

// Address of cell (x, y, z) — each coordinate in [-1, 8] — of population plane `base` (already offset to population q and
// block 0), seen from block `blk` whose info line is spread over the warp (lane i holds word i).
template <typename T>
__device__ __forceinline__ const T* cellOf(const T* __restrict__ base, const uint32_t blk, const uint32_t infoWord, const int x,
                                           const int y, const int z, bool& exists)
{
    const int      fx = (x < 0) ? -1 : (x >= kB ? 1 : 0), fy = (y < 0) ? -1 : (y >= kB ? 1 : 0), fz = (z < 0) ? -1 : (z >= kB ? 1 : 0);
    const uint32_t nb = __shfl_sync(0xffffffffu, infoWord, (fx + 1) + 3 * (fy + 1) + 9 * (fz + 1));
    const uint32_t b = (fx | fy | fz) ? nb : blk;
    exists = b != kNoBlock;
    return base + (int64_t)b * kBlockCells + ((z - fz * kB) * (kB * kB) + (y - fy * kB) * kB + (x - fx * kB));
}

template <class COL, typename T, int VEC>
struct BlockCfg
{
    static constexpr int LPR = kB / VEC;               // lanes per row
    static constexpr int THREADS = kBlockCells / VEC;  // 128 (VEC 4), 256 (VEC 2), 512 (VEC 1)
    static constexpr int VALUE_REGS = COL::Q * VEC * (int)sizeof(T) / 4;
    // resident CTAs per SM asked of ptxas (64 K registers per SM)
    static constexpr int MIN_BLOCKS_ = VALUE_REGS <= 80 ? 512 / THREADS : 256 / THREADS;  // 128 / 255 registers per thread
    static constexpr int MIN_BLOCKS = MIN_BLOCKS_ < 1 ? 1 : MIN_BLOCKS_;
};

// ---- loads of population q: the row (y - c_y, z - c_z) of this block or of the neighbour that holds it, plus, on the
// first / last lane of a row, the one cell of the x-neighbour block a shuffle cannot supply
template <class L, int q, typename T, int VEC>
__device__ __forceinline__ void blockLoad(const T* __restrict__ popIn, const BlockArgs& a, const uint32_t blk, const uint32_t infoWord,
                                          const int tx, const int x0, const int y, const int z, T (&v)[VEC], T& edge)
{
    constexpr int cx = L::c(q, 0), cy = L::c(q, 1), cz = L::c(q, 2);
    constexpr int LPR = kB / VEC;
    const int     ys = y - cy, zs = z - cz;
    const T*      base = popIn + q * a.popPitch;
    bool          ex;
    const T*      p = cellOf<T>(base, blk, infoWord, x0, ys, zs, ex);
    ldPred(p, ex, v);
    edge = T(0);
    if constexpr (cx == 1) {  // cell x0 pulls from x0 - 1: the first lane of a row needs x = -1 of the row (ys, zs)
        bool     exe;
        const T* pe = cellOf<T>(base, blk, infoWord, -1, ys, zs, exe);
        edge = ldPred1(pe, tx == 0 && exe);
    } else if constexpr (cx == -1) {
        bool     exe;
        const T* pe = cellOf<T>(base, blk, infoWord, kB, ys, zs, exe);
        edge = ldPred1(pe, tx == LPR - 1 && exe);
    }
}

template <class L, int q, typename T, int VEC>
__device__ __forceinline__ void blockShift(const int tx, T (&v)[VEC], const T edge)
{
    constexpr int cx = L::c(q, 0);
    constexpr int LPR = kB / VEC;
    if constexpr (cx == 1) {
        T e = __shfl_up_sync(0xffffffffu, v[VEC - 1], 1);
        if (tx == 0)
            e = edge;
#pragma unroll
        for (int i = VEC - 1; i > 0; --i)
            v[i] = v[i - 1];
        v[0] = e;
    } else if constexpr (cx == -1) {
        T e = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (tx == LPR - 1)
            e = edge;
#pragma unroll
        for (int i = 0; i < VEC - 1; ++i)
            v[i] = v[i + 1];
        v[VEC - 1] = e;
    }
}

template <class L, typename T, int VEC, int... Qs>
__device__ __forceinline__ void blockLoadAll(std::integer_sequence<int, Qs...>, const T* __restrict__ popIn, const BlockArgs& a,
                                             const uint32_t blk, const uint32_t infoWord, const int tx, const int x0, const int y,
                                             const int z, T (&f)[L::Q][VEC], T (&edge)[L::Q])
{
    (blockLoad<L, Qs, T, VEC>(popIn, a, blk, infoWord, tx, x0, y, z, f[Qs], edge[Qs]), ...);
}
template <class L, typename T, int VEC, int... Qs>
__device__ __forceinline__ void blockShiftAll(std::integer_sequence<int, Qs...>, const int tx, T (&f)[L::Q][VEC], const T (&edge)[L::Q])
{
    (blockShift<L, Qs, T, VEC>(tx, f[Qs], edge[Qs]), ...);
}

// ---- wall fix-up of one cell: in[q] = f_opp(q)(x) + f_opp(q)(x - c_q)   (LbmTools.h:78-96), block-aware addressing.
// The branch is warp-divergent (only lanes with wall neighbours take it) while cellOf shuffles the info line, so this code
// reads the neighbour ids from the copy of the info line the warp parked in shared memory (`nbr`).
template <class L, typename T, int VEC, int P>
__device__ __forceinline__ void blockFixLoad(const T* __restrict__ popIn, const BlockArgs& a, const uint32_t blk,
                                             const uint32_t* __restrict__ nbr, const uint32_t m, const int x, const int y, const int z,
                                             const int i, T (&f)[L::Q][VEC], T& tb)
{
    constexpr int q = Pairs<L>::lo(P), o = L::opp(q);
    const bool    bq = (m >> q) & 1u, bo = (m >> o) & 1u;
    // bq: f_o(x) + f_o(x - c_q);   bo (only): f_q(x) + f_q(x + c_q).  Predicated volatile PTX loads without a consumer here, the
    // first operand straight into the slot it replaces (see fixLoad in lbm_step.cuh: selects behind plain loads serialised the pairs)
    const int      s = bq ? -1 : 1;
    const int      xn = x + s * L::c(q, 0), yn = y + s * L::c(q, 1), zn = z + s * L::c(q, 2);
    const int      fx = (xn < 0) ? -1 : (xn >= kB ? 1 : 0), fy = (yn < 0) ? -1 : (yn >= kB ? 1 : 0), fz = (zn < 0) ? -1 : (zn >= kB ? 1 : 0);
    const uint32_t bn = (fx | fy | fz) ? nbr[(fx + 1) + 3 * (fy + 1) + 9 * (fz + 1)] : blk;
    const T*       plane = popIn + (bq ? o : q) * a.popPitch;
    const T*       own = plane + (int64_t)blk * kBlockCells + (z * (kB * kB) + y * kB + x);
    const T*       ngh = plane + (int64_t)bn * kBlockCells + ((zn - fz * kB) * (kB * kB) + (yn - fy * kB) * kB + (xn - fx * kB));
    f[q][i] = ldPredKeepNc1(own, bq, f[q][i]);
    f[o][i] = ldPredKeepNc1(own, bo && !bq, f[o][i]);
    tb = ldPred1(ngh, (bq || bo) && bn != kNoBlock);
}
template <class L, typename T, int VEC, int P>
__device__ __forceinline__ void blockFixUse(const T* __restrict__ popIn, const BlockArgs& a, const uint32_t blk, const uint32_t* __restrict__ nbr,
                                            const uint32_t m, const int x, const int y, const int z, const int i, const T tb,
                                            T (&f)[L::Q][VEC])
{
    constexpr int q = Pairs<L>::lo(P), o = L::opp(q);
    const bool    bq = (m >> q) & 1u, bo = (m >> o) & 1u;
    if (bq)
        f[q][i] = f[q][i] + tb;
    else if (bo)
        f[o][i] = f[o][i] + tb;
    if (bq && bo) {  // walls on both sides along c_q: f_q(x) + f_q(x + c_q)
        const int  xn = x + L::c(q, 0), yn = y + L::c(q, 1), zn = z + L::c(q, 2);
        const int  fx = (xn < 0) ? -1 : (xn >= kB ? 1 : 0), fy = (yn < 0) ? -1 : (yn >= kB ? 1 : 0), fz = (zn < 0) ? -1 : (zn >= kB ? 1 : 0);
        const uint32_t bn = (fx | fy | fz) ? nbr[(fx + 1) + 3 * (fy + 1) + 9 * (fz + 1)] : blk;
        const T*   plane = popIn + q * a.popPitch;
        const T    t1 = __ldg(plane + (int64_t)blk * kBlockCells + (z * (kB * kB) + y * kB + x));
        const T    t2 = bn != kNoBlock ? __ldg(plane + (int64_t)bn * kBlockCells + ((zn - fz * kB) * (kB * kB) + (yn - fy * kB) * kB + (xn - fx * kB))) : T(0);
        f[o][i] = t1 + t2;
    }
}
template <class L, typename T, int VEC, int P0, int... Ps>
__device__ __forceinline__ void blockFixChunk(std::integer_sequence<int, Ps...>, const T* __restrict__ popIn, const BlockArgs& a,
                                              const uint32_t blk, const uint32_t* __restrict__ nbr, const uint32_t m, const int x, const int y,
                                              const int z, const int i, T (&f)[L::Q][VEC])
{
    T tb[sizeof...(Ps)];
    (blockFixLoad<L, T, VEC, P0 + Ps>(popIn, a, blk, nbr, m, x, y, z, i, f, tb[Ps]), ...);
    (blockFixUse<L, T, VEC, P0 + Ps>(popIn, a, blk, nbr, m, x, y, z, i, tb[Ps], f), ...);
}
template <class L, typename T, int VEC>
__device__ __forceinline__ void blockFixCell(const T* __restrict__ popIn, const BlockArgs& a, const uint32_t blk, const uint32_t* __restrict__ nbr,
                                             const uint32_t m, const int x, const int y, const int z, const int i, T (&f)[L::Q][VEC])
{
    constexpr int NP = Pairs<L>::N;
    constexpr int CH = sizeof(T) == 4 ? 9 : 5;
    constexpr int N0 = NP < CH ? NP : CH, N1 = NP < 2 * CH ? NP : 2 * CH;
    blockFixChunk<L, T, VEC, 0>(std::make_integer_sequence<int, N0>{}, popIn, a, blk, nbr, m, x, y, z, i, f);
    if constexpr (NP > CH)
        blockFixChunk<L, T, VEC, CH>(std::make_integer_sequence<int, N1 - CH>{}, popIn, a, blk, nbr, m, x, y, z, i, f);
    if constexpr (NP > 2 * CH)
        blockFixChunk<L, T, VEC, 2 * CH>(std::make_integer_sequence<int, NP - 2 * CH>{}, popIn, a, blk, nbr, m, x, y, z, i, f);
    static_assert(NP <= 3 * CH, "chunking covers three batches");
}

// =============================================================== the kernel
template <class COL, typename T, int VEC>
__global__ void __launch_bounds__(BlockCfg<COL, T, VEC>::THREADS, BlockCfg<COL, T, VEC>::MIN_BLOCKS) k_block_step(const BlockArgs a)
{
    constexpr int Q = COL::Q;
    using L = Lattice<Q>;
    using Cfg = BlockCfg<COL, T, VEC>;
    const uint32_t blk = a.firstBlock + blockIdx.x;
    const int      t = threadIdx.x, lane = t & 31;
    const int      tx = t % Cfg::LPR, y = (t / Cfg::LPR) % kB, z = t / (Cfg::LPR * kB);
    const int      x0 = tx * VEC;
    const int64_t  cellOff = (int64_t)blk * kBlockCells + (z * (kB * kB) + y * kB + x0);
    const T*       popIn = reinterpret_cast<const T*>(a.in);

    // the block's info line (one 128-byte line: lane i holds word i) gates every neighbour address; the CTA that ran
    // kPrefetchAhead blocks earlier pulled it into L2, and this one does the same for a later block
    uint32_t infoWord, word0;
    asm volatile("ld.global.nc.u32 %0, [%1];\n" : "=r"(infoWord) : "l"(a.info + (int64_t)blk * 32 + lane));
    // the flag word of the block's first cell carries NLBM_FLAG_BLOCK_PLAIN (set by nlbm_block_wall_mask when every cell of the
    // block is bulk without a wall neighbour): such a block needs no flag words at all — 4 of 156 bytes per cell (ncu r02k: the
    // flag words were three quarters of the kernel's DRAM traffic above the algorithmic bytes).  It travels next to the info
    // line, which gates the population loads anyway, and like it was pulled into L2 by an earlier CTA.
    asm volatile("ld.global.nc.u32 %0, [%1];\n" : "=r"(word0) : "l"(a.flags + (int64_t)blk * kBlockCells));
    if (t == 0 && blockIdx.x + kPrefetchAhead < a.nBlocks) {
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(a.info + ((int64_t)blk + kPrefetchAhead) * 32));
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(a.flags + ((int64_t)blk + kPrefetchAhead) * kBlockCells));
    }
    const bool plainBlock = (word0 & NLBM_FLAG_BLOCK_PLAIN) != 0 && !a.eagerFlags;
    uint32_t   fl[VEC];
    ldFlags<VEC>(a.flags + cellOff, !plainBlock, fl);
    if (plainBlock) {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            fl[i] = kPlainBulk;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        fl[i] &= ~(uint32_t)NLBM_FLAG_BLOCK_PLAIN;  // (the marker is not part of the cell's flags)
    // every population load of the thread goes out before anything is consumed
    T f[Q][VEC], edge[Q];
    blockLoadAll<L, T, VEC>(std::make_integer_sequence<int, Q>{}, popIn, a, blk, infoWord, tx, x0, y, z, f, edge);

    bool plain = true, anyBulk = false, allBulk = true;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        plain = plain && fl[i] == kPlainBulk;
        anyBulk = anyBulk || flagIsBulk(fl[i]);
        allBulk = allBulk && flagIsBulk(fl[i]);
    }
    if (!__any_sync(0xffffffffu, anyBulk))
        return;  // nothing to update in this warp's rows

    blockShiftAll<L, T, VEC>(std::make_integer_sequence<int, Q>{}, tx, f, edge);

    T* out0 = reinterpret_cast<T*>(a.out) + cellOff;
    __shared__ uint32_t sInfo[Cfg::THREADS / 32][32];
    if (__any_sync(0xffffffffu, !plain)) {
        // the divergent fix-up code cannot shuffle: park the info line where single lanes can read it
        sInfo[t >> 5][lane] = infoWord;
        __syncwarp();
        const uint32_t* nbr = sInfo[t >> 5];
        if (!plain) {
            const bool mixed = anyBulk && !allBulk;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const bool keepOld = mixed && !flagIsBulk(fl[i]);
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    f[q][i] = ldPredCoherent1(out0 + q * a.popPitch + i, keepOld, f[q][i]);
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const uint32_t m = fl[i] & kMaskBits;
                if (m != 0 && flagIsBulk(fl[i]))
                    blockFixCell<L, T, VEC>(popIn, a, blk, nbr, m, x0 + i, y, z, i, f);
            }
        }
    }

    const typename COL::Compute omega = (typename COL::Compute)a.omega;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        T p[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            p[q] = f[q][i];
        const bool bulk = flagIsBulk(fl[i]);
        if (!BulkOnly<COL>::value || bulk)
            COL::run(p, omega);
#pragma unroll
        for (int q = 0; q < Q; ++q)
            f[q][i] = bulk ? p[q] : f[q][i];
    }
    if (!anyBulk)
        return;
#pragma unroll
    for (int q = 0; q < Q; ++q)
        stVec<T, VEC>(out0 + q * a.popPitch, f[q]);
}

template <class COL, typename T, int VEC>
inline cudaError_t launchBlockStepVec(const BlockArgs& a, uint32_t nBlocks, cudaStream_t st)
{
    if (nBlocks == 0)
        return cudaSuccess;
    k_block_step<COL, T, VEC><<<nBlocks, BlockCfg<COL, T, VEC>::THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

template <class COL, typename T>
inline cudaError_t launchBlockStep(const BlockArgs& a, uint32_t nBlocks, cudaStream_t st)
{
    // widest access whose values fit 128 registers: D3Q19 -> 16 bytes, D3Q27 -> 8 bytes
    constexpr int maxVec = 16 / (int)sizeof(T);
    constexpr int vec = (COL::Q * maxVec * (int)sizeof(T) / 4 > 80) ? maxVec / 2 : maxVec;
    return launchBlockStepVec<COL, T, vec>(a, nBlocks, st);
}

}  // namespace nlbm
