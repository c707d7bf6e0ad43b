// lbm_step_fast.cu — NLBM_ARITH_FAST instantiations (fused multiply-add in the storage precision; the
// store-float/compute-double kind keeps the reference expressions and only allows contraction).
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

cudaError_t launchStepFast(StepKind kind, const DenseArgs& a, const StepLaunch& l, cudaStream_t st)
{
    switch (kind) {
        case kD3Q19_F32:
            return go<CollideD3Q19Fast<float, 1>, float>(a, l, st);
        case kD3Q19_F64:
            return go<CollideD3Q19Fast<double, 1>, double>(a, l, st);
        case kD3Q19_F32C64:
            return go<CollideD3Q19Ref<float, double, 1>, float>(a, l, st);
        case kD3Q27_F32:
            return go<CollideD3Q27Fast<float, 1>, float>(a, l, st);
        case kD3Q27_F64:
            return go<CollideD3Q27Fast<double, 1>, double>(a, l, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launchMultiFast(StepKind kind, const DenseArgs& a, const MultiArgs& m, const StepLaunch& l, cudaStream_t st)
{
    switch (kind) {
        case kD3Q19_F32:
            return launchMulti<CollideD3Q19Fast<float, 1>, float>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q19_F64:
            return launchMulti<CollideD3Q19Fast<double, 1>, double>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q19_F32C64:
            return launchMulti<CollideD3Q19Ref<float, double, 1>, float>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q27_F32:
            return launchMulti<CollideD3Q27Fast<float, 1>, float>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
        case kD3Q27_F64:
            return launchMulti<CollideD3Q27Fast<double, 1>, double>(a, m, l.nzView, l.vec, l.rowsLog2, l.rpwSel, st);
    }
    return cudaErrorInvalidValue;
}

void tmaTileShape(int elemBytes, int nx, int* tx, int* ty)
{
    int l2;
    if (elemBytes == 4)
        tmaGeometry<CollideD3Q19Fast<float, 1>, float>(nx, &l2, tx, ty);
    else
        tmaGeometry<CollideD3Q19Fast<double, 1>, double>(nx, &l2, tx, ty);
}
}  // namespace nlbm
