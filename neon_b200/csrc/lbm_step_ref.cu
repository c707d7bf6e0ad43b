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
}  // namespace nlbm
