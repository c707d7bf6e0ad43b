// lbm_step_tma.cuh — persistent, TMA-fed variant of the fused pull-stream + BGK kernel (sm_100a).
//
// Why: the direct kernel (lbm_step.cuh) keeps 19 x 16 B per thread in flight only while a warp sits in its load
// phase; ncu shows it latency-bound (long-scoreboard stalls, ~20 % warps active, 55 % DRAM-active cycles).  Here one
// elected producer thread per CTA streams whole tiles into a ring of shared-memory stages with
// cp.async.bulk.tensor (TMA), 4-5 stages (~200 KB) in flight per SM at all times, while two groups of four consumer
// warps alternate over the stages.
//
// The y and z pull shifts cost nothing: population q of the tile at (x0, y0, z) is fetched with the box origin at
// (., y0 - c_qy, z - c_qz) of a 4-D tensor map (x, y, plane, q) over the SoA field, and out-of-box coordinates are
// zero-filled by the TMA unit (no bounds predicates).  The x shift cannot ride on the box origin: TMA needs a 16-byte
// aligned start address (measured on B200: an inner coordinate that is not a multiple of 16 B raises "illegal
// instruction", tools/tma_probe.cu).  Populations with c_x != 0 are therefore fetched through a second map whose box is
// one 16-byte vector wider (origin x0 - PAD for c_x = +1, x0 for c_x = -1) and consumers pick their cells with one
// aligned 16-byte shared-memory load plus one scalar load; populations with c_x = 0 are one aligned vector load.
// Consumers then release the stage and run the same fix-up / collide / store tail as the direct kernel (finishCells).
//
// Ring protocol per stage s: full[s] (1 arrival + transaction bytes) producer -> consumers; empty[s] (4 arrivals: one
// per warp of the consuming group) consumers -> producer; info[s][row] carries the row-summary bits of the tile so that
// consumers know which 32-cell chunks hold bulk / non-plain cells without touching global memory.
// A group only waits on the stages of ITS tiles, so it does not observe every phase of a barrier; a parity wait alone
// could then pass one whole phase early (parity of phase P equals that of P-2).  tileId[s] — written by the producer
// before it arms full[s] — disambiguates: a consumer accepts a stage only when the parity wait passes AND the stage
// carries its tile index (consumerAcquire).
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint in lbm_api.cu)

#include "lbm_step.cuh"

namespace nlbm {

struct TileArgs
{
    int32_t ntx, nty, nz;   // tiles per row / per plane column, planes of the view
    int32_t txLog2;         // tile width TX = 1 << txLog2 (32..128), TY = TILE / TX rows
    int32_t bytesA, bytesB; // shared-memory bytes of one unshifted / one x-shifted population tile (128-byte multiples)
    int32_t txBytes;        // bytes TMA delivers per stage
    int32_t stages, stageBytes;
    int32_t flagOff, flagBytes;  // flag tile inside a stage (fetched only for tiles with non-plain cells)
};

template <class COL, typename T>
struct TmaCfg
{
    static constexpr int Q = COL::Q;
    static constexpr int CPT = 16 / (int)sizeof(T);  // cells per thread = elements of one 16-byte vector = halo pad
    static constexpr int GROUP_THREADS = 128;
    // consumer groups of four warps: three (13 warps with the producer -> 128 registers per thread) when the Q x CPT
    // population values of a thread fit that budget (D3Q19), else two (9 warps -> 168 registers; D3Q27, mixed precision)
    static constexpr int VALUE_REGS = Q * CPT * (int)sizeof(typename COL::Compute) / 4;
    static constexpr int MAX_GROUPS = VALUE_REGS <= 76 ? 3 : 2;
    static constexpr int TILE = GROUP_THREADS * CPT;  // 512 cells (4-byte) / 256 cells (8-byte)
    static constexpr int SMEM_MAX = 227 * 1024;
    static constexpr int TAIL_BYTES = 2048;  // barriers + per-stage row info
    static constexpr int MAX_STAGES = 8;
    static constexpr int MAX_TX = 128;       // box width TX + CPT must stay <= 256
    static constexpr int THREADS = 32 + MAX_GROUPS * GROUP_THREADS;
    static constexpr int MAX_ROWS = 16;
    static constexpr int FIX_CHUNK = 5;  // wall fix-up loads per batch (register budget)
    static_assert(MAX_STAGES * (16 + MAX_ROWS * 4 + 4) <= TAIL_BYTES, "tail too small");
    __host__ __device__ static constexpr int shiftedBefore(int q)
    {
        int n = 0;
        for (int k = 0; k < q; ++k)
            n += Lattice<Q>::c(k, 0) != 0;
        return n;
    }
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void     mbarInit(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NLBM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NLBM_DONE;\n"
        "bra NLBM_WAIT;\n"
        "NLBM_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbarTest(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Wait until stage `bar` holds tile `tile` and all its bytes have landed (see the header comment on phase aliasing).
__device__ __forceinline__ void consumerAcquire(uint32_t bar, uint32_t parity, const volatile uint32_t* tileId, uint32_t tile)
{
    for (;;) {
        mbarWait(bar, parity);
        if (*tileId != tile) {  // passed on the parity of an older phase: the stage still belongs to tile - S
            __nanosleep(64);
            continue;
        }
        if (mbarTest(bar, parity))  // the id was current: re-check that this very phase has completed
            return;
    }
}
__device__ __forceinline__ void mbarArrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbarArriveExpectTx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmaLoad3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmaLoad4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

// shared-memory offset of population q inside a stage
template <class Cfg, int q>
__device__ __forceinline__ int popOffset(const TileArgs& ta)
{
    constexpr int nb = Cfg::shiftedBefore(q);
    return (q - nb) * ta.bytesA + nb * ta.bytesB;
}

template <class Cfg, typename T, int q>
__device__ __forceinline__ void tmaIssueOne(uint32_t stage, const CUtensorMap* mapA, const CUtensorMap* mapB, uint32_t bar,
                                            const TileArgs& ta, int x0, int y0, int zm)
{
    using L = Lattice<Cfg::Q>;
    constexpr int cx = L::c(q, 0);
    // the box origin stays 16-byte aligned in x: c_x = +1 reads [x0 - PAD, x0 + TX), c_x = -1 reads [x0, x0 + TX + PAD)
    tmaLoad4d(stage + popOffset<Cfg, q>(ta), cx == 0 ? mapA : mapB, bar, cx == 1 ? x0 - Cfg::CPT : x0, y0 - L::c(q, 1),
              zm - L::c(q, 2), q);
}
template <class Cfg, typename T, int... Qs>
__device__ __forceinline__ void tmaIssueAll(std::integer_sequence<int, Qs...>, uint32_t stage, const CUtensorMap* mapA,
                                            const CUtensorMap* mapB, uint32_t bar, const TileArgs& ta, int x0, int y0, int zm)
{
    (tmaIssueOne<Cfg, T, Qs>(stage, mapA, mapB, bar, ta, x0, y0, zm), ...);
}

// my CPT cells of population q out of the staged tile: cell x takes the value stored for x - c_x
template <class Cfg, typename T, int q>
__device__ __forceinline__ void pullSmemOne(const unsigned char* stage, const TileArgs& ta, int lx, int ly, T (&f)[Cfg::CPT])
{
    using L = Lattice<Cfg::Q>;
    using V = typename Vec<T, Cfg::CPT>::type;
    constexpr int cx = L::c(q, 0), CPT = Cfg::CPT;
    const int     TX = 1 << ta.txLog2;
    const T*      base = reinterpret_cast<const T*>(stage + popOffset<Cfg, q>(ta));
    if constexpr (cx == 0) {
        const V  v = *reinterpret_cast<const V*>(base + ly * TX + lx);
        const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int c = 0; c < CPT; ++c)
            f[c] = e[c];
    } else if constexpr (cx == 1) {
        const T* row = base + ly * (TX + CPT);  // row[i] holds x0 - CPT + i
        const V  v = *reinterpret_cast<const V*>(row + lx + CPT);
        const T* e = reinterpret_cast<const T*>(&v);
        f[0] = row[lx + CPT - 1];
#pragma unroll
        for (int c = 1; c < CPT; ++c)
            f[c] = e[c - 1];
    } else {
        const T* row = base + ly * (TX + CPT);  // row[i] holds x0 + i
        const V  v = *reinterpret_cast<const V*>(row + lx);
        const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int c = 0; c < CPT - 1; ++c)
            f[c] = e[c + 1];
        f[CPT - 1] = row[lx + CPT];
    }
}
template <class Cfg, typename T, int... Qs>
__device__ __forceinline__ void pullSmemAll(std::integer_sequence<int, Qs...>, const unsigned char* stage, const TileArgs& ta,
                                            int lx, int ly, T (&f)[Cfg::Q][Cfg::CPT])
{
    (pullSmemOne<Cfg, T, Qs>(stage, ta, lx, ly, f[Qs]), ...);
}

// ------------------------------------------------------------------ the kernel
// grid = min(#tiles, #SMs) persistent CTAs; CTA b takes tiles b, b + gridDim.x, ... (x fastest, then y, then plane)
template <class COL, typename T>
__global__ void __launch_bounds__(TmaCfg<COL, T>::THREADS, 1)
    k_dense_step_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmF, const DenseArgs a, const TileArgs ta)
{
    using Cfg = TmaCfg<COL, T>;
    constexpr int Q = Cfg::Q, CPT = Cfg::CPT, TILE = Cfg::TILE;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int      S = ta.stages;
    uint64_t*      bars = reinterpret_cast<uint64_t*>(smem + S * ta.stageBytes);  // full[0..S), empty[0..S)
    uint32_t*      info = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::MAX_STAGES);  // [S][MAX_ROWS]: spec | bulk << 16
    uint32_t*      tileId = info + Cfg::MAX_STAGES * Cfg::MAX_ROWS;                 // [S]: CTA-local index of the staged tile
    const uint32_t fullBar = smemAddr(bars), emptyBar = smemAddr(bars + Cfg::MAX_STAGES);
    const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            tileId[s] = 0xffffffffu;
            mbarInit(fullBar + 8 * s, 1);
            mbarInit(emptyBar + 8 * s, Cfg::GROUP_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int TX = 1 << ta.txLog2, TY = TILE >> ta.txLog2;
    const int tilesPerPlane = ta.ntx * ta.nty;
    const int nTiles = tilesPerPlane * ta.nz;
    const int myTiles = ((int)blockIdx.x < nTiles) ? (nTiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        // ------------------------------------------------ producer warp: lanes fetch the row summaries, lane 0 drives TMA
        const uint32_t cm = (1u << (TX >> 5)) - 1u;
        int            s = 0;
        uint32_t       ph = 0;
        for (int i = 0; i < myTiles; ++i) {
            const int k = blockIdx.x + i * gridDim.x;
            const int tz = k / tilesPerPlane, rem = k - tz * tilesPerPlane;
            const int ty = rem / ta.ntx, tx = rem - ty * ta.ntx;
            const int x0 = tx << ta.txLog2, y0 = ty * TY;
            const int zm = a.zm0 + tz + (tz >= a.fold ? a.skip : 0);
            uint32_t  rowInfo = 0;
            if (lane < TY && y0 + lane < a.ny) {
                const uint2 w = __ldg(a.summary + ((int64_t)zm * a.ny + y0 + lane) * a.wpr + (x0 >> 10));
                const int   sh = (x0 >> 5) & 31;
                rowInfo = ((w.x >> sh) & cm) | (((w.y >> sh) & cm) << 16);
            }
            const bool anyBulk = __any_sync(0xffffffffu, (rowInfo >> 16) != 0);
            const bool anySpec = __any_sync(0xffffffffu, (rowInfo & 0xffffu) != 0);
            mbarWait(emptyBar + 8 * s, ph ^ 1u);  // the stage (and its info row) is free
            if (lane < Cfg::MAX_ROWS)
                info[s * Cfg::MAX_ROWS + lane] = rowInfo;
            if (lane == Cfg::MAX_ROWS)
                tileId[s] = (uint32_t)i;
            __syncwarp();
            if (lane == 0) {
                if (anyBulk) {
                    const uint32_t stage = smemAddr(smem) + s * ta.stageBytes;
                    mbarArriveExpectTx(fullBar + 8 * s, ta.txBytes + (anySpec ? ta.flagBytes : 0));
                    if (anySpec)  // flag words of the tile ride along: consumers never wait on a global flag load
                        tmaLoad3d(stage + ta.flagOff, &tmF, fullBar + 8 * s, x0, y0, zm);
                    tmaIssueAll<Cfg, T>(std::make_integer_sequence<int, Q>{}, stage, &tmA, &tmB, fullBar + 8 * s, ta, x0, y0, zm);
                } else {
                    mbarArrive(fullBar + 8 * s);  // nothing to update in this tile: consumers just hand the stage back
                }
            }
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else {
        // ------------------------------------------------ consumers: group g takes this CTA's tiles g, g+groups, ...
        const int groups = ((int)blockDim.x - 32) / Cfg::GROUP_THREADS;
        const int g = (warp - 1) >> 2;
        const int t = ((warp - 1) & 3) * 32 + lane;  // thread index inside the group
        const int i0 = t * CPT;                      // first of my CPT cells in the tile (x fastest)
        const int lx = i0 & (TX - 1), ly = i0 >> ta.txLog2;
        for (int i = g; i < myTiles; i += groups) {
            const int k = blockIdx.x + i * gridDim.x;
            const int tz = k / tilesPerPlane, rem = k - tz * tilesPerPlane;
            const int ty = rem / ta.ntx, tx = rem - ty * ta.ntx;
            const int x = (tx << ta.txLog2) + lx, y = ty * TY + ly;
            const int zm = a.zm0 + tz + (tz >= a.fold ? a.skip : 0);
            const int s = i % S;
            consumerAcquire(fullBar + 8 * s, (uint32_t)(i / S) & 1u, tileId + s, (uint32_t)i);
            const uint32_t ri = info[s * Cfg::MAX_ROWS + ly];
            const bool     hasBulk = (ri >> (16 + (lx >> 5))) & 1u;
            const bool     special = (ri >> (lx >> 5)) & 1u;
            T              f[Q][CPT];
            uint32_t       fl[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                fl[c] = kPlainBulk;
            if (hasBulk) {
                const unsigned char* stage = smem + s * ta.stageBytes;
                pullSmemAll<Cfg, T>(std::make_integer_sequence<int, Q>{}, stage, ta, lx, ly, f);
                if (special) {
                    using FV = typename Vec<float, CPT>::type;
                    const FV        w = *reinterpret_cast<const FV*>(stage + ta.flagOff + i0 * 4);
                    const uint32_t* e = reinterpret_cast<const uint32_t*>(&w);
#pragma unroll
                    for (int c = 0; c < CPT; ++c)
                        fl[c] = e[c];
                }
            }
            __syncwarp();
            if (lane == 0)
                mbarArrive(emptyBar + 8 * s);  // my warp's reads of the stage are done
            if (!hasBulk)
                continue;
            const int64_t cellOff = (int64_t)zm * a.pitch_z + (int64_t)y * a.pitch_y + x;
            finishCells<COL, T, CPT, Cfg::FIX_CHUNK>(a, reinterpret_cast<const T*>(a.in) + cellOff, reinterpret_cast<T*>(a.out) + cellOff, fl,
                                     special, f);
        }
    }
}

// ------------------------------------------------------------------ host side
// Tile geometry for a row length nx: TX = smallest power of two >= nx, clamped to [32, MAX_TX].
template <class COL, typename T>
inline void tmaGeometry(int nx, int* txLog2, int* tx, int* ty)
{
    using Cfg = TmaCfg<COL, T>;
    int w = 32, l = 5;
    while (w < nx && w < Cfg::MAX_TX && w < Cfg::TILE) {
        w *= 2;
        ++l;
    }
    *txLog2 = l;
    *tx = w;
    *ty = Cfg::TILE / w;
}

template <class COL, typename T>
inline cudaError_t launchStepTma(const DenseArgs& a, int nzView, const void* tmapA, const void* tmapB, const void* tmapF,
                                 int groups, int numSms, cudaStream_t st)
{
    using Cfg = TmaCfg<COL, T>;
    using L = Lattice<Cfg::Q>;
    if (nzView <= 0)
        return cudaSuccess;
    TileArgs ta;
    int      TX, TY;
    tmaGeometry<COL, T>(a.nx, &ta.txLog2, &TX, &TY);
    const int nShift = Cfg::shiftedBefore(Cfg::Q), nPlain = Cfg::Q - nShift;
    const int rawB = (TX + Cfg::CPT) * TY * (int)sizeof(T);
    ta.bytesA = Cfg::TILE * (int)sizeof(T);
    ta.bytesB = (rawB + 127) / 128 * 128;
    ta.txBytes = nPlain * ta.bytesA + nShift * rawB;
    ta.flagOff = nPlain * ta.bytesA + nShift * ta.bytesB;
    ta.flagBytes = Cfg::TILE * 4;
    ta.stageBytes = ta.flagOff + (ta.flagBytes + 127) / 128 * 128;
    ta.stages = (Cfg::SMEM_MAX - Cfg::TAIL_BYTES) / ta.stageBytes;
    if (ta.stages > Cfg::MAX_STAGES)
        ta.stages = Cfg::MAX_STAGES;
    if (ta.stages < 2)
        return cudaErrorInvalidConfiguration;
    const int smemBytes = ta.stages * ta.stageBytes + Cfg::TAIL_BYTES;
    (void)sizeof(L);

    static unsigned char configured[64] = {0};  // per instantiation and device (benign race: the call is idempotent)
    auto                 kern = k_dense_step_tma<COL, T>;
    int                  dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_MAX);
        if (e != cudaSuccess)
            return e;
        if (dev >= 0 && dev < 64)
            configured[dev] = 1;
    }
    ta.ntx = (a.nx + TX - 1) / TX;
    ta.nty = (a.ny + TY - 1) / TY;
    ta.nz = nzView;
    const long long tiles = (long long)ta.ntx * ta.nty * ta.nz;
    if (tiles > 0x7fffffffLL)
        return cudaErrorInvalidValue;
    const int grid = (int)(tiles < numSms ? tiles : numSms);
    if (groups < 1 || groups > Cfg::MAX_GROUPS)
        groups = Cfg::MAX_GROUPS;
    kern<<<grid, 32 + groups * Cfg::GROUP_THREADS, smemBytes, st>>>(*reinterpret_cast<const CUtensorMap*>(tmapA),
                                                                     *reinterpret_cast<const CUtensorMap*>(tmapB),
                                                                     *reinterpret_cast<const CUtensorMap*>(tmapF), a, ta);
    return cudaGetLastError();
}

}  // namespace nlbm
