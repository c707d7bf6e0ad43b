// lbm_block_fast.cu — NLBM_ARITH_FAST instantiations of the block-sparse step kernel.
#include "lbm_block.cuh"
#include "lbm_host.h"

namespace nlbm {
cudaError_t launchBlockStepFast(StepKind kind, const BlockArgs& a, uint32_t nBlocks, cudaStream_t st)
{
    switch (kind) {
        case kD3Q19_F32:
            return launchBlockStep<CollideD3Q19Fast<float, 1>, float>(a, nBlocks, st);
        case kD3Q19_F64:
            return launchBlockStep<CollideD3Q19Fast<double, 1>, double>(a, nBlocks, st);
        case kD3Q19_F32C64:
            return launchBlockStep<CollideD3Q19Ref<float, double, 1>, float>(a, nBlocks, st);
        case kD3Q27_F32:
            return launchBlockStep<CollideD3Q27Fast<float, 1>, float>(a, nBlocks, st);
        case kD3Q27_F64:
            return launchBlockStep<CollideD3Q27Fast<double, 1>, double>(a, nBlocks, st);
    }
    return cudaErrorInvalidValue;
}
}  // namespace nlbm
