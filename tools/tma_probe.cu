// tools/tma_probe.cu — stand-alone probe of cp.async.bulk.tensor configurations (development aid, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tools/tma_probe.cu && for i in 0 1 2 3 4 5 6 7; do ./tma_probe $i; done
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int c0, int c1, int c2, int c3, int boxElems, int dstOff)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t  b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(smem) + dstOff;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(boxElems * 4) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
                         "l"(&m), "r"(c0), "r"(c1), "r"(b)
                         : "memory");
        else if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
                         "l"(&m), "r"(c0), "r"(c1), "r"(c2), "r"(b)
                         : "memory");
        else
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(d),
                "l"(&m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(b)
                : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b)
        : "memory");
    for (int i = threadIdx.x; i < boxElems; i += blockDim.x)
        out[i] = reinterpret_cast<float*>(smem + dstOff)[i];
}

int main(int argc, char** argv)
{
    int            cfg = argc > 1 ? atoi(argv[1]) : 0;
    EncodeTiledFn  enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    void*          p = nullptr;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc = (EncodeTiledFn)p;
    const int nx = 64, ny = 8, nzm = 8, Q = 19, py = 128;
    size_t    n = (size_t)py * ny * nzm * Q;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i)
        h[i] = (float)i;
    float *dptr, *out;
    cudaMalloc(&dptr, n * 4);
    cudaMalloc(&out, 65536);
    cudaMemcpy(dptr, h.data(), n * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t  dims[4] = {nx, ny, nzm, Q};
    cuuint64_t  str[3] = {py * 4, (cuuint64_t)py * ny * 4, (cuuint64_t)py * ny * nzm * 4};
    cuuint32_t  box[4] = {64, 8, 1, 1};
    cuuint32_t  es[4] = {1, 1, 1, 1};
    int         rank = 4, c[4] = {1, 0, 1, 0}, dstOff = 0;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    switch (cfg) {
        case 0: rank = 2; dims[0] = py; dims[1] = ny * nzm * Q; c[0] = 0; c[1] = 0; break;
        case 1: break;
        case 2: c[0] = c[2] = 0; break;
        case 3: l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE; break;
        case 4: dims[0] = py; break;
        case 5: rank = 3; dims[2] = nzm * Q; c[3] = 0; break;
        case 6: c[0] = -1; c[1] = -1; break;
        case 7: box[0] = 32; box[1] = 16; break;
        case 8: c[0] = 0; c[1] = 1; c[2] = 1; c[3] = 3; break;          // only the inner coordinate needs alignment?
        case 9: c[0] = -4; c[1] = -1; c[2] = -1; c[3] = 0; break;       // aligned negative start, zero fill
        case 10: c[0] = 4; c[1] = 0; c[2] = 0; c[3] = 0; break;
        case 11: box[0] = 68; box[1] = 4; c[0] = -4; c[1] = 2; c[2] = 3; c[3] = 18; break;  // box wider than the row, not /32
        case 13: c[0] = 0; c[2] = 0; dstOff = 16; break;                // shared-memory destination only 16-byte aligned
        case 14: c[0] = 0; c[2] = 0; dstOff = 64; break;
        case 12: c[0] = 2; c[1] = 0; c[2] = 0; c[3] = 0; break;         // 8-byte aligned only
    }
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, dptr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int boxElems = box[0] * box[1];
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    if (rank == 2)
        k<2><<<1, 128, 65536 + 64>>>(m, out, c[0], c[1], c[2], c[3], boxElems, dstOff);
    else if (rank == 3)
        k<3><<<1, 128, 65536 + 64>>>(m, out, c[0], c[1], c[2], c[3], boxElems, dstOff);
    else
        k<4><<<1, 128, 65536 + 64>>>(m, out, c[0], c[1], c[2], c[3], boxElems, dstOff);
    cudaError_t e = cudaDeviceSynchronize();
    float       o[4] = {0, 0, 0, 0};
    if (e == cudaSuccess)
        cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
    printf("cfg %d: encode=%d kernel=%s out=%g %g %g %g\n", cfg, (int)r, cudaGetErrorName(e), o[0], o[1], o[2], o[3]);
    return 0;
}
