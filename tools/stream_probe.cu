// tools/stream_probe.cu — what HBM bandwidth can the LBM ACCESS PATTERN reach on this GPU, independent of the arithmetic?
// (development aid, not product code).  Every mode moves 19 SoA planes of nx*ny*nz floats from `in` to `out` with the
// D3Q19 pull shifts in y and z, i.e. exactly the 152 B/cell of the fused collide-and-stream kernel, and reports GB/s.
//
//   mode 0  flat float4 copy of the same byte count (the MEASURED_PEAKS-style ceiling)
//   mode 1  one thread = 4 cells, 19 x LDG.128 -> 19 x STG.128 (the "direct" kernel's pattern), 256-thread blocks
//   mode 2  persistent: TMA tile loads -> shared memory -> TMA tile stores, no SM data path at all (pure DMA)
//   mode 3  persistent: TMA tile loads -> LDS.128 -> STG.128 by consumer warps          (the current TMA kernel's pattern)
//   mode 4  persistent: TMA tile loads -> LDS.128 -> (group barrier) -> STS.128 in place -> TMA tile stores
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/stream_probe tools/stream_probe.cu
//   tools/stream_probe [nz=256] [reps=10]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "../neon_b200/csrc/lbm_step.cuh"  // CollideD3Q19Fast: the product's arithmetic, for the +compute variants

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int Q = 19;
using Col = nlbm::CollideD3Q19Fast<float, 1>;
__device__ __forceinline__ void collide4(float4 (&f)[Q], float omega)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float p[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            p[q] = reinterpret_cast<float*>(&f[q])[i];
        Col::run(p, omega);
#pragma unroll
        for (int q = 0; q < Q; ++q)
            reinterpret_cast<float*>(&f[q])[i] = p[q];
    }
}
__constant__ int c_cy[Q] = {0, -1, 0, -1, 1, 0, 0, -1, -1, 0, 0, 1, 0, 1, -1, 0, 0, 1, 1};
__constant__ int c_cz[Q] = {0, 0, -1, 0, 0, -1, 1, -1, 1, 0, 0, 0, 1, 0, 0, 1, -1, 1, -1};

struct Dim
{
    int     nx, ny, nz;
    int64_t pz, pq;
};

// ------------------------------------------------------------------ mode 0
__global__ void k_flat(const float4* __restrict__ in, float4* __restrict__ out, size_t n4)
{
    size_t       i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride), d = __ldg(in + i + 3 * stride);
        __stcs(out + i, a);
        __stcs(out + i + stride, b);
        __stcs(out + i + 2 * stride, c);
        __stcs(out + i + 3 * stride, d);
    }
    for (; i < n4; i += stride)
        __stcs(out + i, __ldg(in + i));
}

// ------------------------------------------------------------------ mode 1
template <int MINB, bool COMPUTE>
__global__ void __launch_bounds__(256, MINB) k_direct(const float* __restrict__ in, float* __restrict__ out, Dim d, float omega)
{
    const int x = (blockIdx.x * blockDim.y + threadIdx.y) * 128 + threadIdx.x * 4;
    const int y = blockIdx.y * blockDim.z + threadIdx.z;
    const int z = blockIdx.z;
    if (x >= d.nx || y >= d.ny)
        return;
    float4 f[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int ys = y - c_cy[q], zs = z - c_cz[q];
        ys = min(max(ys, 0), d.ny - 1);
        zs = min(max(zs, 0), d.nz - 1);
        f[q] = __ldg(reinterpret_cast<const float4*>(in + q * d.pq + zs * d.pz + (int64_t)ys * d.nx + x));
    }
    if (COMPUTE)
        collide4(f, omega);
#pragma unroll
    for (int q = 0; q < Q; ++q)
        __stcs(reinterpret_cast<float4*>(out + q * d.pq + z * d.pz + (int64_t)y * d.nx + x), f[q]);
}

// ------------------------------------------------------------------ TMA helpers
__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void     mbarInit(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void     mbarWait(uint32_t b, uint32_t ph)
{
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DL;\nbra WL;\nDL:\n}\n" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbarExpect(uint32_t b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmaLoad(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmaStore(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(src), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

struct TArgs
{
    int ntx, nty, nz, txLog2, ty, stages, groups, mode, compute;
};
constexpr int TILE = 512, TILE_BYTES = TILE * 4, STAGE_BYTES = Q * TILE_BYTES, MAXS = 8;

// warp 0: load producer; warp 1: store issuer (modes 2, 4); warps 2..: consumer groups of 4 warps (modes 3, 4)
__global__ void __launch_bounds__(64 + 3 * 128, 1)
    k_tma(const __grid_constant__ CUtensorMap mIn, const __grid_constant__ CUtensorMap mOut, const float* __restrict__ in,
          float* __restrict__ out, Dim d, TArgs t)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int      S = t.stages;
    uint64_t*      bars = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
    const uint32_t full = sa(bars), empty = sa(bars + MAXS), ready = sa(bars + 2 * MAXS);
    const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbarInit(full + 8 * s, 1);
            mbarInit(empty + 8 * s, t.mode == 3 ? 4 : 1);  // mode 3: consumers free the stage; 2/4: the store issuer does
            mbarInit(ready + 8 * s, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int TX = 1 << t.txLog2;
    const int perPlane = t.ntx * t.nty, nTiles = perPlane * t.nz;
    const int my = ((int)blockIdx.x < nTiles) ? (nTiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (warp == 0) {
        if (lane == 0) {
            int      s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < my; ++i) {
                const int k = blockIdx.x + i * gridDim.x;
                const int tz = k / perPlane, rem = k - tz * perPlane, ty = rem / t.ntx, tx = rem - ty * t.ntx;
                mbarWait(empty + 8 * s, ph ^ 1u);
                mbarExpect(full + 8 * s, STAGE_BYTES);
                const uint32_t st = sa(smem) + s * STAGE_BYTES;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    tmaLoad(st + q * TILE_BYTES, &mIn, full + 8 * s, tx * TX, ty * t.ty - c_cy[q], tz - c_cz[q], q);
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && (t.mode == 2 || t.mode == 4)) {
            int      s = 0;
            uint32_t ph = 0;
            int      prev = -1;
            for (int i = 0; i < my; ++i) {
                const int k = blockIdx.x + i * gridDim.x;
                const int tz = k / perPlane, rem = k - tz * perPlane, ty = rem / t.ntx, tx = rem - ty * t.ntx;
                mbarWait((t.mode == 2 ? full : ready) + 8 * s, ph);
                const uint32_t st = sa(smem) + s * STAGE_BYTES;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    tmaStore(&mOut, st + q * TILE_BYTES, tx * TX, ty * t.ty, tz, q);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (prev >= 0) {  // the previous tile's stores have read their stage: hand it back
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    mbarArrive(empty + 8 * prev);
                }
                prev = s;
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (prev >= 0)
                mbarArrive(empty + 8 * prev);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (t.mode >= 3) {
        const int g = (warp - 2) >> 2;
        if (g >= t.groups)
            return;
        const int tid = ((warp - 2) & 3) * 32 + lane;
        const int i0 = tid * 4, lx = i0 & (TX - 1), ly = i0 >> t.txLog2;
        for (int i = g; i < my; i += t.groups) {
            const int k = blockIdx.x + i * gridDim.x;
            const int tz = k / perPlane, rem = k - tz * perPlane, ty = rem / t.ntx, tx = rem - ty * t.ntx;
            const int s = i % S;
            mbarWait(full + 8 * s, (uint32_t)(i / S) & 1u);
            float4*       st = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
            float4        f[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q)
                f[q] = st[q * (TILE / 4) + tid];
            if (t.mode == 3) {
                __syncwarp();
                if (lane == 0)
                    mbarArrive(empty + 8 * s);
                if (t.compute)
                    collide4(f, 1.2f);
                const int64_t off = (int64_t)tz * d.pz + (int64_t)(ty * t.ty + ly) * d.nx + tx * TX + lx;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    __stcs(reinterpret_cast<float4*>(out + q * d.pq + off), f[q]);
            } else {
                asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");  // every read of the stage is done
                if (t.compute)
                    collide4(f, 1.2f);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    f[q].x += 1.0f;
                    st[q * (TILE / 4) + tid] = f[q];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    mbarArrive(ready + 8 * s);
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int nz = argc > 1 ? atoi(argv[1]) : 256, reps = argc > 2 ? atoi(argv[2]) : 10;
    Dim       d{512, 512, nz, 512 * 512, (int64_t)512 * 512 * nz};
    CK(cudaFree(0));
    void*                           p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const size_t  n = (size_t)Q * d.pq;
    float *       in, *out;
    CK(cudaMalloc(&in, n * 4));
    CK(cudaMalloc(&out, n * 4));
    {
        std::vector<float> h(1 << 20, 0.05f);
        for (size_t o = 0; o < n; o += h.size())
            CK(cudaMemcpy(in + o, h.data(), std::min(h.size(), n - o) * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(out, 0, n * 4));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double gb = 2.0 * n * 4 / 1e9;
    auto         timeit = [&](const char* name, auto launch) {
        launch();
        launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r)
            launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        printf("%-58s %8.3f ms  %8.1f GB/s  (%.0f MLUPS-equivalent)\n", name, ms, gb / (ms * 1e-3), (double)d.pq / (ms * 1e3));
        fflush(stdout);
    };

    for (int blocks : {sms * 4, sms * 8, sms * 16, sms * 32})
        for (int thr : {256, 512}) {
            char nm[96];
            snprintf(nm, sizeof nm, "mode0 flat float4 copy, %d blocks x %d", blocks, thr);
            timeit(nm, [&] { k_flat<<<blocks, thr>>>((const float4*)in, (float4*)out, n / 4); });
        }
    {
        dim3 block(32, 4, 2), grid(1, d.ny / 2, d.nz);
        timeit("mode1 direct 19xLDG.128->19xSTG.128, minBlocks 2", [&] { k_direct<2, false><<<grid, block>>>(in, out, d, 1.2f); });
        timeit("mode1 direct, minBlocks 3 (<=80 regs)", [&] { k_direct<3, false><<<grid, block>>>(in, out, d, 1.2f); });
        dim3 block2(32, 1, 8), grid2(4, d.ny / 8, d.nz);
        timeit("mode1 direct, block = 8 rows x 1 segment, minBlocks 2", [&] { k_direct<2, false><<<grid2, block2>>>(in, out, d, 1.2f); });
        timeit("mode1 direct + BGK collide, minBlocks 2", [&] { k_direct<2, true><<<grid, block>>>(in, out, d, 1.2f); });
        timeit("mode1 direct + BGK collide, minBlocks 3", [&] { k_direct<3, true><<<grid, block>>>(in, out, d, 1.2f); });
        timeit("mode1 direct + BGK collide, 8 rows x 1 segment", [&] { k_direct<2, true><<<grid2, block2>>>(in, out, d, 1.2f); });
    }
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int txLog2 : {7, 8}) {
        const int   TX = 1 << txLog2, TY = TILE / TX;
        CUtensorMap mIn, mOut;
        cuuint64_t  dims[4] = {(cuuint64_t)d.nx, (cuuint64_t)d.ny, (cuuint64_t)d.nz, Q};
        cuuint64_t  str[3] = {(cuuint64_t)d.nx * 4, (cuuint64_t)d.pz * 4, (cuuint64_t)d.pq * 4};
        cuuint32_t  box[4] = {(cuuint32_t)TX, (cuuint32_t)TY, 1, 1}, es[4] = {1, 1, 1, 1};
        for (int promo : {0}) {
            CUtensorMapL2promotion l2 = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
            if (enc(&mIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
                enc(&mOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("tensor map encode failed\n");
                return 1;
            }
            for (int compute : {0, 1})
            for (int mode : {2, 3, 4})
                for (int stages : {3, 4, 5, 6})
                    for (int groups : {2, 3}) {
                        if (compute && mode == 2)
                            continue;
                        if (mode == 2 && groups != 2)
                            continue;
                        // a group must see every phase of the stages it waits on: stages % groups == 0 (the product kernel
                        // lifts this with a per-stage tile tag)
                        if (mode != 2 && (stages % groups != 0 || stages * STAGE_BYTES > 220 * 1024))
                            continue;
                        TArgs t{d.nx / TX, d.ny / TY, d.nz, txLog2, TY, stages, groups, mode, compute};
                        char  nm[96];
                        snprintf(nm, sizeof nm, "mode%d%s TMA box %dx%d l2promo %s, %d stages, %d groups", mode, compute ? "+BGK" : "", TX, TY,
                                 promo ? "256B" : "none", stages, groups);
                        const int smemBytes = stages * STAGE_BYTES + 3 * MAXS * 8;
                        timeit(nm, [&] { k_tma<<<sms, 64 + 128 * groups, smemBytes>>>(mIn, mOut, in, out, d, t); });
                    }
        }
    }
    return 0;
}
