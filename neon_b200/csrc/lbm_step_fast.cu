// lbm_step_fast.cu — NLBM_ARITH_FAST instantiations (fused multiply-add in the storage precision; the
// store-float/compute-double kind keeps the reference expressions and only allows contraction).
#include "lbm_host.h"
#include "lbm_step.cuh"

namespace nlbm {
cudaError_t launchStepFast(StepKind kind, const DenseArgs& a, int nzView, int vec, int rowsLog2, cudaStream_t st)
{
    switch (kind) {
        case kD3Q19_F32:
            return launchStep<CollideD3Q19Fast<float, 1>, float>(a, nzView, vec, rowsLog2, st);
        case kD3Q19_F64:
            return launchStep<CollideD3Q19Fast<double, 1>, double>(a, nzView, vec, rowsLog2, st);
        case kD3Q19_F32C64:
            return launchStep<CollideD3Q19Ref<float, double, 1>, float>(a, nzView, vec, rowsLog2, st);
        case kD3Q27_F32:
            return launchStep<CollideD3Q27Fast<float, 1>, float>(a, nzView, vec, rowsLog2, st);
        case kD3Q27_F64:
            return launchStep<CollideD3Q27Fast<double, 1>, double>(a, nzView, vec, rowsLog2, st);
    }
    return cudaErrorInvalidValue;
}
}  // namespace nlbm
