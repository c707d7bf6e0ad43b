// lbm_step_ref.cu — NLBM_ARITH_REFERENCE instantiations.  Compiled with -fmad=false: no contraction, IEEE division,
// so every operation rounds as the reference's CPU build (x86-64, -O2) does.
#include "lbm_host.h"
#include "lbm_step_tma.cuh"

namespace nlbm {
namespace {
template <class COL, typename T>
cudaError_t go(const DenseArgs& a, const StepLaunch& l, cudaStream_t st)
{
    if (l.tmapA)
        return launchStepTma<COL, T>(a, l.nzView, l.tmapA, l.tmapB, l.tmapF, l.groups, l.numSms, st);
    return launchStep<COL, T>(a, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
}
}  // namespace

cudaError_t launchStepRef(StepKind kind, const DenseArgs& a, const StepLaunch& l, cudaStream_t st)
{
    switch (kind) {
        case kD3Q19_F32:
            if (l.exact)
                return go<CollideD3Q19Exact<0>, float>(a, l, st);
            return go<CollideD3Q19Ref<float, float, 0>, float>(a, l, st);
        case kD3Q19_F64:
            return go<CollideD3Q19Ref<double, double, 0>, double>(a, l, st);
        case kD3Q19_F32C64:
            return go<CollideD3Q19Ref<float, double, 0>, float>(a, l, st);
        case kD3Q27_F32:
            return go<CollideD3Q27Ref<float, 0>, float>(a, l, st);
        case kD3Q27_F64:
            return go<CollideD3Q27Ref<double, 0>, double>(a, l, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launchMultiRef(StepKind kind, const DenseArgs& a, const MultiArgs& m, const StepLaunch& l, cudaStream_t st)
{
    switch (kind) {
        case kD3Q19_F32:
            if (l.exact)
                return launchMulti<CollideD3Q19Exact<0>, float>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
            return launchMulti<CollideD3Q19Ref<float, float, 0>, float>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q19_F64:
            return launchMulti<CollideD3Q19Ref<double, double, 0>, double>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q19_F32C64:
            return launchMulti<CollideD3Q19Ref<float, double, 0>, float>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q27_F32:
            return launchMulti<CollideD3Q27Ref<float, 0>, float>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q27_F64:
            return launchMulti<CollideD3Q27Ref<double, 0>, double>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------- self-test of the exact building blocks (tests/)
// kind 0: exact::widenPos(f) against the conversion instruction for EVERY positive normal float (n and seed ignored);
// kind 1: exact::div3 against IEEE division on n pseudo-random (a0, a1, a2, b) inside div3's guard, zero numerators included.
__global__ void k_selftest_exact(int kind, unsigned long long n, unsigned long long seed, unsigned long long* bad)
{
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long       mism = 0;
    if (kind == 0) {
        for (unsigned long long b = 0x00800000ull + tid; b < 0x7f800000ull; b += nthreads) {
            const float f = __uint_as_float((uint32_t)b);
            mism += __double_as_longlong(exact::widenPos(f)) != __double_as_longlong((double)f);
        }
    } else {
        unsigned long long s = seed * 0x9E3779B97F4A7C15ull + tid * 0xD1B54A32D192ED03ull + 1;
        auto               next = [&]() {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            return s;
        };
        for (unsigned long long i = tid; i < n; i += nthreads) {
            const unsigned long long r0 = next(), r1 = next();
            // b: any mantissa, exponent in [-60, 60]; numerators: |a| = b * (random in (2^-k, 1]), random sign, some exactly 0
            const uint32_t eb = 127u - 60u + (uint32_t)((r0 >> 40) % 121u);
            const float    b = __uint_as_float((eb << 23) | (uint32_t)(r0 & 0x7fffffu));
            float          a[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const unsigned long long r = k == 0 ? r1 : next();
                const uint32_t           down = (uint32_t)((r >> 32) % 40u);
                float                    v = b * __uint_as_float(((127u - down) << 23) | (uint32_t)(r & 0x7fffffu)) * 0.5f;
                if (fabsf(v) < 8.673617380e-19f || ((r >> 60) == 0))
                    v = 0.f;
                a[k] = (r >> 59) & 1 ? -v : v;
                if (a[k] == 0.f)
                    a[k] = 0.f;  // +0, as x - x gives
            }
            float q0, q1, q2;
            exact::div3(a[0], a[1], a[2], b, q0, q1, q2);
            mism += __float_as_uint(q0) != __float_as_uint(__fdiv_rn(a[0], b));
            mism += __float_as_uint(q1) != __float_as_uint(__fdiv_rn(a[1], b));
            mism += __float_as_uint(q2) != __float_as_uint(__fdiv_rn(a[2], b));
        }
    }
    if (mism)
        atomicAdd(bad, mism);
}

cudaError_t launchSelftestExact(int kind, unsigned long long n, unsigned long long seed, unsigned long long* dBad, cudaStream_t st)
{
    k_selftest_exact<<<148 * 8, 256, 0, st>>>(kind, n, seed, dBad);
    return cudaGetLastError();
}
}  // namespace nlbm
