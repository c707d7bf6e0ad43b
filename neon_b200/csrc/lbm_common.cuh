// lbm_common.cuh — shared device/host helpers of the sm_100a LBM kernels.
//
// Lattice tables follow the reference numbering exactly:
//   D3Q19: benchmarks/lbm-lid-driven-cavity-flow/src/D3Q19.h:23-44 (velocities), :112-132 (weights);
//          opposite of q is q+10 / q-10, rest population is 9.
//   D3Q27: apps/lbmMultiRes/lattice.h:15-77 (rest population is 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/neon_lbm.h"

namespace nlbm {

// ---------------------------------------------------------------- lattices
template <int Q>
struct Lattice;

template <>
struct Lattice<19>
{
    static constexpr int Q = 19;
    static constexpr int REST = 9;
    __host__ __device__ static constexpr int c(int q, int d)
    {
        constexpr int t[19][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1},
                                  {0, -1, -1}, {0, -1, 1}, {0, 0, 0},  {1, 0, 0},  {0, 1, 0},   {0, 0, 1},  {1, 1, 0},
                                  {1, -1, 0},  {1, 0, 1},  {1, 0, -1}, {0, 1, 1},  {0, 1, -1}};
        return t[q][d];
    }
    __host__ __device__ static constexpr int opp(int q) { return q == 9 ? 9 : (q < 9 ? q + 10 : q - 10); }
    __host__ __device__ static constexpr double w(int q)
    {
        return q == 9 ? 1. / 3. : (((q % 10) < 3) ? 1. / 18. : 1. / 36.);
    }
};

template <>
struct Lattice<27>
{
    static constexpr int Q = 27;
    static constexpr int REST = 0;
    __host__ __device__ static constexpr int c(int q, int d)
    {
        constexpr int t[27][3] = {{0, 0, 0},   {0, 0, -1},   {0, 0, 1},   {0, -1, 0},  {0, -1, -1}, {0, -1, 1}, {0, 1, 0},
                                  {0, 1, -1},  {0, 1, 1},    {-1, 0, 0},  {-1, 0, -1}, {-1, 0, 1},  {-1, -1, 0}, {-1, -1, -1},
                                  {-1, -1, 1}, {-1, 1, 0},   {-1, 1, -1}, {-1, 1, 1},  {1, 0, 0},   {1, 0, -1}, {1, 0, 1},
                                  {1, -1, 0},  {1, -1, -1},  {1, -1, 1},  {1, 1, 0},   {1, 1, -1},  {1, 1, 1}};
        return t[q][d];
    }
    __host__ __device__ static constexpr int opp(int q)
    {
        constexpr int t[27] = {0,  2,  1,  6,  8,  7,  3,  5,  4,  18, 20, 19, 24, 26,
                               25, 21, 23, 22, 9,  11, 10, 15, 17, 16, 12, 14, 13};
        return t[q];
    }
    __host__ __device__ static constexpr double w(int q)
    {
        const int n = (c(q, 0) != 0) + (c(q, 1) != 0) + (c(q, 2) != 0);
        return n == 0 ? 8.0 / 27.0 : (n == 1 ? 2.0 / 27.0 : (n == 2 ? 1.0 / 54.0 : 1.0 / 216.0));
    }
};

// ---------------------------------------------------------------- flag words
constexpr uint32_t kMaskBits = NLBM_FLAG_MASK_BITS;
constexpr uint32_t kPlainBulk = (uint32_t)NLBM_BULK << NLBM_FLAG_CLASS_SHIFT;  // bulk, no wall neighbour
__host__ __device__ inline uint32_t flagClass(uint32_t f) { return (f >> NLBM_FLAG_CLASS_SHIFT) & 3u; }
__host__ __device__ inline bool     flagIsBulk(uint32_t f) { return flagClass(f) == (uint32_t)NLBM_BULK; }

// Row summary that lives behind the per-cell flag words (same buffer): for every row (zm, y) and every
// group of 32 chunks (chunk = 32 consecutive cells in x) one uint2 {special, bulk}:
//   bit i of .x set <=> chunk i holds a cell that is not "plain bulk" (non-bulk, wall bits set, or x >= nx)
//   bit i of .y set <=> chunk i holds at least one bulk cell
// The step kernels skip flag loads and the fix-up code for chunks whose special bit is clear.
constexpr int kChunk = 32;
__host__ __device__ inline int64_t summaryWordsPerRow(int64_t pitch_y) { return (pitch_y / kChunk + 31) / 32; }
__host__ __device__ inline int64_t alignUp(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int64_t flagCellWords(const nlbm_dense_desc& d)
{
    return alignUp((int64_t)(d.nz_local + 2 * d.z_halo) * d.pitch_z, 32);
}
inline __host__ __device__ const uint2* summaryPtr(const nlbm_dense_desc& d)
{
    return reinterpret_cast<const uint2*>(d.flags + flagCellWords(d));
}
// Cell map, behind the row summary in the same buffer: ONE BYTE per 4 consecutive cells of a row (pitch_y / 4 bytes per
// row, rows as in the flag array):  bit i (0..3) set <=> cell 4g+i is bulk;  bit 4+i set <=> cell 4g+i is not "plain bulk"
// (non-bulk, wall bits set, or x >= nx).  A thread of the step kernel reads its byte together with the populations and
// needs the 4-byte flag words only where a bulk cell has wall bits or sits next to a non-bulk cell in the same thread:
// 0.25 B/cell of traffic instead of 4 B/cell, and no dependent round trip for plain threads.
__host__ __device__ inline int64_t cellMapBytesPerRow(int64_t pitch_y) { return pitch_y / 4; }
__host__ __device__ inline int64_t flagRows(const nlbm_dense_desc& d) { return (int64_t)d.ny * (d.nz_local + 2 * d.z_halo); }
inline __host__ __device__ const uint8_t* cellMapPtr(const nlbm_dense_desc& d)
{
    return reinterpret_cast<const uint8_t*>(summaryPtr(d) + flagRows(d) * summaryWordsPerRow(d.pitch_y));
}
__host__ __device__ inline int64_t flagBufferBytes(const nlbm_dense_desc& d)
{
    return (flagCellWords(d) + 2 * flagRows(d) * summaryWordsPerRow(d.pitch_y)) * 4 + alignUp(flagRows(d) * cellMapBytesPerRow(d.pitch_y), 128);
}

enum FlagMode { kFlagWords = 0, kFlagSummaryFirst = 1, kFlagCellMap = 2 };

// ---------------------------------------------------------------- kernel arguments (by value)
struct DenseArgs
{
    const void* in;
    void*       out;
    const uint32_t* flags;
    const uint2*    summary;
    int32_t nx, ny, nzm;  // nzm = nz_local + 2*z_halo memory planes
    int32_t pitch_y;      // elements
    int64_t pitch_z, pitch_q;
    int32_t wpr;          // summary words per row
    int32_t zm0;          // first memory plane of the view
    int32_t fold, skip;   // view planes >= fold are shifted by skip (BOUNDARY view: two slabs)
    int32_t lprLog2;      // direct kernel: log2 of the lanes a warp spends on one row (32 = whole-row warps)
    int32_t flagMode;     // direct kernel: how a thread learns about its cells — kFlagWords: the flag words travel with the
                          // populations; kFlagSummaryFirst: row summary first, flag words only for non-plain chunks;
                          // kFlagCellMap: one byte per 4 cells with the populations, flag words only for non-plain threads
    const uint8_t* cellMap;
    int32_t specXFix;     // direct kernel: fetch the wall fix-up operands of the cells next to the x faces speculatively (cp.async)
    int32_t prefetchXFaces;  // direct kernel: fetch the output-field values of the cells at x = 0 and x = nx-1 speculatively (cp.async)
    const void* keepCache;   // x-face cache of the output field (nlbm_dense_wall_cache_build) or null
    int32_t experiment;   // MEASUREMENT ONLY (results are wrong): 1 every cell is plain bulk, no flag loads; 2 flags loaded but ignored
    double  omega;
    // ---- fused face push (nlbm_dense_step_push): the kernel stores the crossing populations of its two z-boundary planes
    // straight into the z-neighbours' ghost planes and signals them; all null / 0 for the plain step
    int32_t   peerMode;       // 1: boundary planes first, peer stores, signalling
    int32_t   nzLocal;        // planes of the partition (peerMode: plane order 0, nz-1, 1, 2, ...)
    void*     peer[2];        // [0] the neighbour BELOW (receives my plane 0), [1] ABOVE (receives my plane nz-1); may be null
    int64_t   peerOff[2];     // element offset of the ghost plane inside the neighbour's population 0
    int64_t   peerPitchQ[2];  // the neighbour's pitch_q
    uint32_t* peerFlag[2];    // word in the neighbour's memory that receives signalValue once the whole plane is there
    uint32_t* counter;        // 2 words of local memory (zero between launches): warps of plane 0 / plane nz-1 that finished
    uint32_t  signalValue, warpsPerFace;
};

// nlbm_dense_step_n: the second field of the two-field scheme and how the chain of launches is issued (lbm_step.cuh)
struct MultiArgs
{
    const void* fieldB;      // the second field: iteration t reads (t even ? a.in : fieldB) and writes the other one
    const void* keepCacheA;  // x-face cache of a.in (a.keepCache is a.out's, i.e. fieldB's); may be null like a.keepCache
    int32_t     iterations;
    int32_t     chainEarly;  // planes that start on the plane counters (0: one chip-load of blocks, < 0: all)
};

constexpr int kChainPlanesApi = 4096;  // planes a launch chain has counters for (lbm_step.cuh: kChainPlanes)
// second argument of the chained step kernel (k_dense_chain, lbm_step.cuh): one launch per iteration, launched while its
// predecessor still runs (programmatic dependent launch); a tile starts as soon as the planes it reads are complete
struct ChainArgs
{
    unsigned* planeDone;  // one counter per z plane of the view: tiles that finished that plane, summed over the chain's iterations
    unsigned  target;     // a tile of plane z starts once planeDone[z-1], [z], [z+1] >= target (0: first iteration, no wait)
    unsigned  early;      // planes [0, early) start on the counters, the others behind griddepcontrol.wait; planes [0, early] publish
};

// ---------------------------------------------------------------- vector access
template <typename T, int VEC>
struct Vec;
template <>
struct Vec<float, 1>
{
    using type = float;
};
template <>
struct Vec<float, 2>
{
    using type = float2;
};
template <>
struct Vec<float, 4>
{
    using type = float4;
};
template <>
struct Vec<double, 1>
{
    using type = double;
};
template <>
struct Vec<double, 2>
{
    using type = double2;
};

// Predicated, branch-free, non-coherent global loads, written as volatile inline PTX on purpose: the compiler keeps
// volatile asm statements in program order and cannot wrap them in branches, so a kernel that lists all its loads first
// really has all of them in flight before the first use (ncu showed ptxas otherwise interleaving the loads with the
// shuffles that consume them: one exposed DRAM round trip per population).  pred == false yields zeros.
__device__ __forceinline__ void ldPred(const float* p, bool pred, float (&v)[4])
{
    asm volatile(
        "{\n.reg .pred q;\nsetp.ne.u32 q, %5, 0;\nmov.f32 %0, 0f00000000;\nmov.f32 %1, 0f00000000;\nmov.f32 %2, 0f00000000;\n"
        "mov.f32 %3, 0f00000000;\n@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n}\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
        : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPred(const float* p, bool pred, float (&v)[2])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\nmov.f32 %0, 0f00000000;\nmov.f32 %1, 0f00000000;\n"
                 "@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n}\n"
                 : "=f"(v[0]), "=f"(v[1])
                 : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPred(const float* p, bool pred, float (&v)[1])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\nmov.f32 %0, 0f00000000;\n@q ld.global.nc.f32 %0, [%1];\n}\n"
                 : "=f"(v[0])
                 : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPred(const double* p, bool pred, double (&v)[2])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\nmov.f64 %0, 0d0000000000000000;\nmov.f64 %1, 0d0000000000000000;\n"
                 "@q ld.global.nc.v2.f64 {%0, %1}, [%2];\n}\n"
                 : "=d"(v[0]), "=d"(v[1])
                 : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPred(const double* p, bool pred, double (&v)[1])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\nmov.f64 %0, 0d0000000000000000;\n@q ld.global.nc.f64 %0, [%1];\n}\n"
                 : "=d"(v[0])
                 : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ float ldPred1(const float* p, bool pred)
{
    float v[1];
    ldPred(p, pred, v);
    return v[0];
}
__device__ __forceinline__ double ldPred1(const double* p, bool pred)
{
    double v[1];
    ldPred(p, pred, v);
    return v[0];
}
// The same loads as ordinary (coherent) global loads: for the multi-iteration kernel, where the field a thread reads was
// written by other SMs earlier in the SAME launch.  ld.global.nc is undefined for data written during the kernel; an ordinary
// load is ordered by the grid-wide barrier between iterations (which also invalidates L1).  (ld.global.cg compiles to
// LDG.STRONG.GPU on sm_100a and halved the throughput of the tile loop: profiles/r02j_small_sweep.log.)
__device__ __forceinline__ void ldPredCg(const float* p, bool pred, float (&v)[4])
{
    asm volatile(
        "{\n.reg .pred q;\nsetp.ne.u32 q, %5, 0;\nmov.f32 %0, 0f00000000;\nmov.f32 %1, 0f00000000;\nmov.f32 %2, 0f00000000;\n"
        "mov.f32 %3, 0f00000000;\n@q ld.global.v4.f32 {%0, %1, %2, %3}, [%4];\n}\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
        : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPredCg(const float* p, bool pred, float (&v)[2])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\nmov.f32 %0, 0f00000000;\nmov.f32 %1, 0f00000000;\n"
                 "@q ld.global.v2.f32 {%0, %1}, [%2];\n}\n"
                 : "=f"(v[0]), "=f"(v[1])
                 : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPredCg(const float* p, bool pred, float (&v)[1])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\nmov.f32 %0, 0f00000000;\n@q ld.global.f32 %0, [%1];\n}\n"
                 : "=f"(v[0])
                 : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPredCg(const double* p, bool pred, double (&v)[2])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\nmov.f64 %0, 0d0000000000000000;\nmov.f64 %1, 0d0000000000000000;\n"
                 "@q ld.global.v2.f64 {%0, %1}, [%2];\n}\n"
                 : "=d"(v[0]), "=d"(v[1])
                 : "l"(p), "r"((uint32_t)pred));
}
__device__ __forceinline__ void ldPredCg(const double* p, bool pred, double (&v)[1])
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\nmov.f64 %0, 0d0000000000000000;\n@q ld.global.f64 %0, [%1];\n}\n"
                 : "=d"(v[0])
                 : "l"(p), "r"((uint32_t)pred));
}
template <bool COH, typename T, int N>
__device__ __forceinline__ void ldPredSel(const T* p, bool pred, T (&v)[N])
{
    if constexpr (COH)
        ldPredCg(p, pred, v);
    else
        ldPred(p, pred, v);
}
template <bool COH, typename T>
__device__ __forceinline__ T ldPredSel1(const T* p, bool pred)
{
    T v[1];
    ldPredSel<COH>(p, pred, v);
    return v[0];
}
// coherent variants (the output field: cells this kernel never writes, but next to cells it does)
__device__ __forceinline__ float ldPredCoherent1(const float* p, bool pred, float keep)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.f32 %0, [%1];\n}\n" : "+f"(keep) : "l"(p), "r"((uint32_t)pred));
    return keep;
}
__device__ __forceinline__ double ldPredCoherent1(const double* p, bool pred, double keep)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.f64 %0, [%1];\n}\n" : "+d"(keep) : "l"(p), "r"((uint32_t)pred));
    return keep;
}
// read-only-path variants that leave `keep` untouched when the predicate is false: the loaded value lands in the register
// it replaces, so no select sits between the load and its (much later) consumer
__device__ __forceinline__ float ldPredKeepNc1(const float* p, bool pred, float keep)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.nc.f32 %0, [%1];\n}\n" : "+f"(keep) : "l"(p), "r"((uint32_t)pred));
    return keep;
}
__device__ __forceinline__ double ldPredKeepNc1(const double* p, bool pred, double keep)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.nc.f64 %0, [%1];\n}\n" : "+d"(keep) : "l"(p), "r"((uint32_t)pred));
    return keep;
}
__device__ __forceinline__ float ldPredKeepCg1(const float* p, bool pred, float keep)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.f32 %0, [%1];\n}\n" : "+f"(keep) : "l"(p), "r"((uint32_t)pred));
    return keep;
}
__device__ __forceinline__ double ldPredKeepCg1(const double* p, bool pred, double keep)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.f64 %0, [%1];\n}\n" : "+d"(keep) : "l"(p), "r"((uint32_t)pred));
    return keep;
}
template <bool COH, typename T>
__device__ __forceinline__ T ldPredKeepSel1(const T* p, bool pred, T keep)
{
    if constexpr (COH)
        return ldPredKeepCg1(p, pred, keep);
    else
        return ldPredKeepNc1(p, pred, keep);
}
// asynchronous copy of one element from global to shared memory: no destination register, nobody waits until cpAsyncWait
__device__ __forceinline__ void cpAsync1(float* smem, const float* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsync1(double* smem, const double* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// flag words of VEC cells; pred == false yields "undefined" cells (never updated)
template <int VEC>
__device__ __forceinline__ void ldFlags(const uint32_t* p, bool pred, uint32_t (&v)[VEC])
{
    constexpr uint32_t U = (uint32_t)NLBM_UNDEFINED << NLBM_FLAG_CLASS_SHIFT;
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        v[i] = U;
    if constexpr (VEC == 4)
        asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %5, 0;\n@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n}\n"
                     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3])
                     : "l"(p), "r"((uint32_t)pred));
    else if constexpr (VEC == 2)
        asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\n@q ld.global.nc.v2.u32 {%0, %1}, [%2];\n}\n"
                     : "+r"(v[0]), "+r"(v[1])
                     : "l"(p), "r"((uint32_t)pred));
    else
        asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.nc.u32 %0, [%1];\n}\n" : "+r"(v[0]) : "l"(p), "r"((uint32_t)pred));
}
// one byte of the cell map; pred == false yields 0 (no bulk cell)
__device__ __forceinline__ uint32_t ldPredU8(const uint8_t* p, bool pred)
{
    uint32_t v = 0;
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.nc.u8 %0, [%1];\n}\n" : "+r"(v) : "l"(p), "r"((uint32_t)pred));
    return v;
}
__device__ __forceinline__ uint2 ldPredU2(const uint2* p)
{
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

template <typename T, int VEC>
__device__ __forceinline__ void ldVec(const T* __restrict__ p, T (&v)[VEC])
{
    using V = typename Vec<T, VEC>::type;
    const V t = __ldg(reinterpret_cast<const V*>(p));
    const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        v[i] = e[i];
}
// one element, streaming, under a predicate (threads that own bulk AND non-bulk cells store their bulk cells one by one)
__device__ __forceinline__ void stPred1(float* p, const float v, const bool pred)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.global.cs.f32 [%0], %1;\n}\n" ::"l"(p), "f"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void stPred1(double* p, const double v, const bool pred)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.global.cs.f64 [%0], %1;\n}\n" ::"l"(p), "d"(v), "r"((uint32_t)pred) : "memory");
}
template <typename T, int VEC>
__device__ __forceinline__ void stVec(T* __restrict__ p, const T (&v)[VEC])
{
    using V = typename Vec<T, VEC>::type;
    V  t;
    T* e = reinterpret_cast<T*>(&t);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        e[i] = v[i];
    __stcs(reinterpret_cast<V*>(p), t);  // streaming store: not re-read before the next iteration
}

}  // namespace nlbm
