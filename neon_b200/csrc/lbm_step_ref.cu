// lbm_step_ref.cu — NLBM_ARITH_REFERENCE instantiations.  Compiled with -fmad=false: no contraction, IEEE division,
// so every operation rounds as the reference's CPU build (x86-64, -O2) does.
#include "lbm_host.h"
#include "lbm_step.cuh"

namespace nlbm {
cudaError_t launchStepRef(StepKind kind, const DenseArgs& a, int nzView, int vec, int rowsLog2, cudaStream_t st)
{
    switch (kind) {
        case kD3Q19_F32:
            return launchStep<CollideD3Q19Ref<float, float, 0>, float>(a, nzView, vec, rowsLog2, st);
        case kD3Q19_F64:
            return launchStep<CollideD3Q19Ref<double, double, 0>, double>(a, nzView, vec, rowsLog2, st);
        case kD3Q19_F32C64:
            return launchStep<CollideD3Q19Ref<float, double, 0>, float>(a, nzView, vec, rowsLog2, st);
        case kD3Q27_F32:
            return launchStep<CollideD3Q27Ref<float, 0>, float>(a, nzView, vec, rowsLog2, st);
        case kD3Q27_F64:
            return launchStep<CollideD3Q27Ref<double, 0>, double>(a, nzView, vec, rowsLog2, st);
    }
    return cudaErrorInvalidValue;
}
}  // namespace nlbm
