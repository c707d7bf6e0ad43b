// Design tool (not part of the product): instruction throughput of the pipes the bit-exact (REFERENCE arithmetic) collision
// leans on — FP64 add/mul/fma, float<->double conversions, and the integer sequence that widens a float without F2F.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu && /tmp/pipe_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, ILP = 8;

template <int OP>
__global__ void k(float* out, float seed)
{
    float  f[ILP];
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        f[i] = seed + i + threadIdx.x * 1e-3f;
        d[i] = f[i];
    }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0)
                d[i] = __dadd_rn(d[i], 1.25);
            else if (OP == 1)
                d[i] = __dmul_rn(d[i], 1.0000001);
            else if (OP == 2)
                d[i] = __fma_rn(d[i], 1.0000001, 0.5);
            else if (OP == 3) {  // widen + narrow round trip (2 conversions)
                d[i] = (double)f[i];
                asm volatile("" : "+d"(d[i]));
                f[i] = (float)d[i];
                asm volatile("" : "+f"(f[i]));
            } else if (OP == 4) {  // widen only (+1 cheap fp32 op to keep a dependency chain)
                d[i] = (double)f[i];
                asm volatile("" : "+d"(d[i]));
                f[i] = f[i] + __int_as_float(__double2hiint(d[i]) & 1);
            } else if (OP == 5) {  // narrow only
                f[i] = (float)d[i];
                asm volatile("" : "+f"(f[i]));
                d[i] = __hiloint2double(__double2hiint(d[i]), __float_as_int(f[i]) & 0xff);
            } else if (OP == 6) {  // integer widening of a normal float
                const uint32_t b = __float_as_uint(f[i]);
                const uint32_t hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
                const uint32_t lo = b << 29;
                d[i] = __hiloint2double(hi, lo);
                asm volatile("" : "+d"(d[i]));
                f[i] = f[i] + __int_as_float(__double2hiint(d[i]) & 1);
            } else if (OP == 7)
                f[i] = __fmaf_rn(f[i], 1.0000001f, 0.5f);
            else if (OP == 8) {  // Veltkamp rounding of a double to 24 bits: 3 DP ops
                const double t = __dmul_rn(d[i], 536870913.0);
                const double e = __dsub_rn(t, d[i]);
                d[i] = __dsub_rn(t, e);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        s += f[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double opsPerInner)
{
    int             dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float* out;
    const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<OP><<<blocks, threads>>>(out, 1.f);
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(out, 1.f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * threads * ITER * ILP * opsPerInner;
    printf("%-28s %8.3f ms  %8.1f Gop/s  %6.1f op/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / (khz * 1e3),
           khz / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("DADD", 1);
    run<1>("DMUL", 1);
    run<2>("DFMA", 1);
    run<3>("F2F widen+narrow (2 conv)", 2);
    run<4>("F2F widen (+FADD)", 1);
    run<5>("F2F narrow", 1);
    run<6>("integer widen (+FADD)", 1);
    run<7>("FFMA", 1);
    run<8>("Veltkamp round (3 DP)", 3);
    return 0;
}
