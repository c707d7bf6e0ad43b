// lbm_halo.cu — halo update of dense z-slab partitions.
//
// Replaces dField::newHaloUpdate / initHaloUpdateTable (libNeonDomain/.../dGrid/dField_imp.h:341-421,548-641) executed by
// DataTransferContainer::run (libNeonSet/include/Neon/set/container/DataTransferContainer.h:38-55): the reference issues
// one cudaMemcpyPeerAsync per population per direction behind a host-blocking stream sync.  Here ONE launch moves the
// planes of the populations that actually cross the face (c_z == dir).  Every population plane of a z-slice is a
// contiguous run of pitch_z elements, so the copy is a pure 16-byte streaming copy; with dst mapped from a peer GPU the
// stores travel over NVLink.
#include "lbm_common.cuh"
#include "lbm_host.h"

namespace nlbm {

// Copy kernels of the halo update: few, fat blocks on purpose (1024 threads, four 16-byte copies in flight per thread, a handful of
// blocks per plane).  They run at high priority NEXT TO the INTERNAL step kernel, whose blocks own half an SM's registers each —
// every resident block of a copy kernel, however small, keeps one of them off the chip for as long as it lives.  A few thousand
// short blocks of 256 threads (round 2, first half) cost the INTERNAL kernel ~25-45 us per iteration; the stores are bound by the
// NVLink round trip, not by how many SMs issue them, and the faces have a whole iteration to arrive.
constexpr int kPushThreads = 1024, kPushUnroll = 4, kPushBlocksPerPlane = 4, kCopyBlocksPerPlane = 8;

__global__ void __launch_bounds__(kPushThreads) k_plane_copy(const char* __restrict__ src, char* __restrict__ dst, const PlaneList pl,
                                                             const size_t vecPerPlane)
{
    const int    p = blockIdx.y;
    const uint4* s = reinterpret_cast<const uint4*>(src + pl.src[p]);
    uint4*       d = reinterpret_cast<uint4*>(dst + pl.dst[p]);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t       i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kPushUnroll - 1) * stride < vecPerPlane; i += kPushUnroll * stride) {
        uint4 v[kPushUnroll];
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k)
            v[k] = __ldcs(s + i + k * stride);
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k)
            d[i + k * stride] = v[k];
    }
    for (; i < vecPerPlane; i += stride)
        d[i] = __ldcs(s + i);
}

cudaError_t launchPlaneCopy(const void* src, void* dst, const PlaneList& pl, size_t planeBytes, cudaStream_t st)
{
    if (pl.n == 0 || planeBytes == 0)
        return cudaSuccess;
    const size_t vecs = planeBytes / 16;
    size_t       bx = (vecs + kPushThreads * kPushUnroll - 1) / (kPushThreads * kPushUnroll);
    if (bx > (size_t)kCopyBlocksPerPlane)
        bx = kCopyBlocksPerPlane;
    if (bx == 0)
        bx = 1;
    dim3 grid((unsigned)bx, pl.n);
    k_plane_copy<<<grid, kPushThreads, 0, st>>>((const char*)src, (char*)dst, pl, vecs);
    return cudaGetLastError();
}

// Both faces of a partition in ONE launch, signalling included: planes [0, nUp) go to the neighbour above, the rest to the
// neighbour below; the last block to finish publishes `value` in both neighbours' flag words (what k_flag_signal would do
// in two more launches).  counter: one word of local memory, zero between launches.
struct Push2Args
{
    PlaneList pl;
    int       nUp;
    char*     dst[2];       // [0] neighbour above, [1] neighbour below (peer mappings)
    uint32_t* flag[2];      // their flag words (peer mappings), null where there is no neighbour
    uint32_t* counter;
    uint32_t  value, blocks;
};
// (few fat blocks: see k_plane_copy)
__global__ void __launch_bounds__(kPushThreads) k_face_push2(const char* __restrict__ src, const Push2Args a, const size_t vecPerPlane)
{
    const int    p = blockIdx.y;
    const uint4* s = reinterpret_cast<const uint4*>(src + a.pl.src[p]);
    uint4*       d = reinterpret_cast<uint4*>(a.dst[p < a.nUp ? 0 : 1] + a.pl.dst[p]);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t       i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kPushUnroll - 1) * stride < vecPerPlane; i += kPushUnroll * stride) {
        uint4 v[kPushUnroll];
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k)
            v[k] = __ldcs(s + i + k * stride);
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k)
            d[i + k * stride] = v[k];
    }
    for (; i < vecPerPlane; i += stride)
        d[i] = __ldcs(s + i);
    __threadfence_system();  // my peer stores are visible system-wide before my block counts as done
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(a.counter, 1u);
        if (done == a.blocks - 1) {
            *a.counter = 0;  // ready for the next launch
            __threadfence_system();
            for (int k = 0; k < 2; ++k)
                if (a.flag[k])
                    *reinterpret_cast<volatile uint32_t*>(a.flag[k]) = a.value;
            __threadfence_system();
        }
    }
}

cudaError_t launchFacePush2(const void* src, const PlaneList& pl, int nUp, void* dstUp, void* dstDown, uint32_t* flagUp,
                            uint32_t* flagDown, uint32_t* counter, uint32_t value, size_t planeBytes, cudaStream_t st)
{
    if (pl.n == 0 || planeBytes == 0)
        return cudaSuccess;
    const size_t vecs = planeBytes / 16;
    size_t       bx = (vecs + kPushThreads * kPushUnroll - 1) / (kPushThreads * kPushUnroll);
    if (bx > (size_t)kPushBlocksPerPlane)
        bx = kPushBlocksPerPlane;
    if (bx == 0)
        bx = 1;
    Push2Args a;
    a.pl = pl;
    a.nUp = nUp;
    a.dst[0] = (char*)dstUp;
    a.dst[1] = (char*)dstDown;
    a.flag[0] = flagUp;
    a.flag[1] = flagDown;
    a.counter = counter;
    a.value = value;
    a.blocks = (uint32_t)(bx * pl.n);
    dim3 grid((unsigned)bx, pl.n);
    k_face_push2<<<grid, kPushThreads, 0, st>>>((const char*)src, a, vecs);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- block-sparse faces
// One z-slice (64 cells = 256 / 512 contiguous bytes) of every boundary block and crossing population.
struct SliceArgs
{
    int      comps[27];
    int      ncomps;
    int64_t  srcPopPitch, dstPopPitch;  // bytes
    uint32_t srcFirst, dstFirst, nBlocks;
    int      sliceOff, vecPerSlice;     // byte offset of the slice inside a block tile, 16-byte vectors per slice
    int      blockBytes;
};
// (few fat blocks, several copies in flight per thread: see k_face_push2)
__global__ void __launch_bounds__(kPushThreads) k_block_slice_copy(const char* __restrict__ src, char* __restrict__ dst, const SliceArgs s)
{
    const int64_t total = (int64_t)s.nBlocks * s.vecPerSlice;
    const int     c = blockIdx.y;
    const char*   sp = src + s.comps[c] * s.srcPopPitch + (int64_t)s.srcFirst * s.blockBytes + s.sliceOff;
    char*         dp = dst + s.comps[c] * s.dstPopPitch + (int64_t)s.dstFirst * s.blockBytes + s.sliceOff;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto          offsetOf = [&](int64_t i) {
        const int64_t b = i / s.vecPerSlice, v = i - b * s.vecPerSlice;
        return b * s.blockBytes + v * 16;
    };
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kPushUnroll - 1) * stride < total; i += kPushUnroll * stride) {
        uint4   v[kPushUnroll];
        int64_t o[kPushUnroll];
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k) {
            o[k] = offsetOf(i + k * stride);
            v[k] = __ldcs(reinterpret_cast<const uint4*>(sp + o[k]));
        }
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k)
            *reinterpret_cast<uint4*>(dp + o[k]) = v[k];
    }
    for (; i < total; i += stride) {
        const int64_t o = offsetOf(i);
        *reinterpret_cast<uint4*>(dp + o) = __ldcs(reinterpret_cast<const uint4*>(sp + o));
    }
}

cudaError_t launchBlockSliceCopy(const void* src, void* dst, int elemBytes, const int* comps, int ncomps, int64_t srcPopPitch,
                                 int64_t dstPopPitch, uint32_t srcFirst, uint32_t dstFirst, uint32_t nBlocks, int zSlice, cudaStream_t st)
{
    if (nBlocks == 0 || ncomps == 0)
        return cudaSuccess;
    SliceArgs s;
    for (int i = 0; i < ncomps; ++i)
        s.comps[i] = comps[i];
    s.ncomps = ncomps;
    s.srcPopPitch = srcPopPitch * elemBytes;
    s.dstPopPitch = dstPopPitch * elemBytes;
    s.srcFirst = srcFirst;
    s.dstFirst = dstFirst;
    s.nBlocks = nBlocks;
    s.blockBytes = 512 * elemBytes;
    s.sliceOff = zSlice * 64 * elemBytes;
    s.vecPerSlice = 64 * elemBytes / 16;
    const int64_t total = (int64_t)nBlocks * s.vecPerSlice;
    int64_t       bx = (total + kPushThreads * kPushUnroll - 1) / (kPushThreads * kPushUnroll);
    if (bx > kPushBlocksPerPlane)
        bx = kPushBlocksPerPlane;
    if (bx < 1)
        bx = 1;
    dim3 grid((unsigned)bx, ncomps);
    k_block_slice_copy<<<grid, kPushThreads, 0, st>>>((const char*)src, (char*)dst, s);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- device-side ordering between GPUs
// One process per GPU cannot order a neighbour's stream with CUDA events without a host hand-shake per iteration.  The
// peer-store halo transport therefore orders with flag words: after its face copy (stream order) the sender publishes a
// counter in the receiver's memory, and the receiver's stream holds a one-thread kernel that waits for it.  The wait
// gives up after timeoutMs so that a lost neighbour cannot hang the GPU — and then it TRAPS: the step kernel behind it
// would read a stale or half-written ghost plane, so the error must not be survivable.  *err is incremented first (for a
// host that still can read it); after the trap every later CUDA call of the process fails.
__global__ void k_flag_signal(volatile uint32_t* flag, uint32_t value)
{
    __threadfence_system();  // the face copy of the previous kernel in this stream is visible system-wide first
    *flag = value;
    __threadfence_system();
}

// waits for up to two flag words (a partition has at most two z-neighbours); a null pointer is not waited for
__global__ void k_flag_wait(const volatile uint32_t* flag0, const volatile uint32_t* flag1, uint32_t value, unsigned long long timeoutNs,
                            int32_t* err)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const volatile uint32_t* flags[2] = {flag0, flag1};
    for (int k = 0; k < 2; ++k) {
        if (!flags[k])
            continue;
        while ((int32_t)(*flags[k] - value) < 0) {  // counters only grow; the signed difference tolerates wrap-around
            __nanosleep(200);
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeoutNs) {
                if (err) {
                    atomicAdd(err, 1);
                    __threadfence_system();
                }
                __trap();  // fatal: whatever follows in this stream would compute on a ghost plane that never arrived
            }
        }
    }
    __threadfence_system();
}

cudaError_t launchFlagSignal(uint32_t* flag, uint32_t value, cudaStream_t st)
{
    k_flag_signal<<<1, 1, 0, st>>>(flag, value);
    return cudaGetLastError();
}
cudaError_t launchFlagWait(const uint32_t* flag0, const uint32_t* flag1, uint32_t value, uint32_t timeoutMs, int32_t* err, cudaStream_t st)
{
    k_flag_wait<<<1, 1, 0, st>>>(flag0, flag1, value, (unsigned long long)timeoutMs * 1000000ull, err);
    return cudaGetLastError();
}

}  // namespace nlbm
