// lbm_host.h — internal declarations shared by the translation units of libneon_lbm.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/neon_lbm.h"

namespace nlbm {

enum StepKind { kD3Q19_F32 = 0, kD3Q19_F64 = 1, kD3Q19_F32C64 = 2, kD3Q27_F32 = 3, kD3Q27_F64 = 4 };

struct DenseArgs;
// lbm_step_ref.cu (-fmad=false) / lbm_step_fast.cu
// tmapA != nullptr selects the TMA-fed persistent kernel (lbm_step_tma.cuh), else the direct kernel (lbm_step.cuh)
struct StepLaunch
{
    int         nzView, vec, rowsLog2;
    int         rpwSel;  // direct kernel: rows per warp selector (0 default, else log2(rows) + 1)
    const void* tmapA;  // CUtensorMap over pop_in with box (TX, TY), or nullptr
    const void* tmapB;  // same with box (TX + 16 B, TY): populations with c_x != 0
    const void* tmapF;  // 3-D map over the flag words, box (TX, TY, 1)
    int         groups; // consumer groups of the TMA kernel (0 = default)
    int         numSms;
    int         exact;  // REFERENCE arithmetic through the conversion-lean evaluation where one exists (same bits)
};
cudaError_t launchStepRef(StepKind kind, const DenseArgs& a, const StepLaunch& l, cudaStream_t st);
cudaError_t launchStepFast(StepKind kind, const DenseArgs& a, const StepLaunch& l, cudaStream_t st);
// nlbm_dense_step_n: several iterations as a chain of dependent launches (direct kernel only); m names the second field
struct MultiArgs;
cudaError_t launchMultiRef(StepKind kind, const DenseArgs& a, const MultiArgs& m, const StepLaunch& l, cudaStream_t st);
cudaError_t launchMultiFast(StepKind kind, const DenseArgs& a, const MultiArgs& m, const StepLaunch& l, cudaStream_t st);
cudaError_t launchSelftestExact(int kind, unsigned long long n, unsigned long long seed, unsigned long long* dBad, cudaStream_t st);
// tile width / rows of the TMA kernel for a population of elemBytes and a row of nx cells
void tmaTileShape(int elemBytes, int nx, int* tx, int* ty);

// lbm_setup.cu
cudaError_t launchSummary(const nlbm_dense_desc& d, cudaStream_t st);
cudaError_t launchWallCacheBuild(const nlbm_dense_desc& d, int q, int elemBytes, cudaStream_t st);
cudaError_t launchFlagsFromClasses(const nlbm_dense_desc& d, const uint8_t* cls, int zmFirst, int nPlanes, cudaStream_t st);
cudaError_t launchClassify(const nlbm_dense_desc& d, int geom, const double* sphere, cudaStream_t st);
cudaError_t launchWallMask(const nlbm_dense_desc& d, int q, int32_t* d_bad, cudaStream_t st);
template <typename S>
cudaError_t launchInitPop(const nlbm_dense_desc& d, int q, double ulb, cudaStream_t st);
template <typename S, typename C>
cudaError_t launchRhoU(const nlbm_dense_desc& d, void* rho, void* u, cudaStream_t st);

// lbm_block_ref.cu / lbm_block_fast.cu
struct BlockArgs;
cudaError_t launchBlockStepRef(StepKind kind, const BlockArgs& a, uint32_t nBlocks, cudaStream_t st, bool exact = false);
cudaError_t launchBlockStepFast(StepKind kind, const BlockArgs& a, uint32_t nBlocks, cudaStream_t st);
cudaError_t launchBlockClassify(const nlbm_block_desc& d, int geom, const double* sphere, const uint32_t* activeMask, cudaStream_t st);
cudaError_t launchBlockWallMask(const nlbm_block_desc& d, int q, int32_t* bad, cudaStream_t st);
template <typename S>
cudaError_t launchBlockInitPop(const nlbm_block_desc& d, int q, double ulb, cudaStream_t st);

// lbm_halo.cu
cudaError_t launchBlockSliceCopy(const void* src, void* dst, int elemBytes, const int* comps, int ncomps, int64_t srcPopPitch,
                                 int64_t dstPopPitch, uint32_t srcFirst, uint32_t dstFirst, uint32_t nBlocks, int zSlice, cudaStream_t st);
struct PlaneList
{
    int     n;
    int64_t src[54];  // byte offsets (up to 2 x 27: both faces of a partition in one launch)
    int64_t dst[54];
};
cudaError_t launchFlagSignal(uint32_t* flag, uint32_t value, cudaStream_t st);
cudaError_t launchFlagWait(const uint32_t* flag0, const uint32_t* flag1, uint32_t value, uint32_t timeoutMs, int32_t* err, cudaStream_t st);
cudaError_t launchFacePush2(const void* src, const PlaneList& pl, int nUp, void* dstUp, void* dstDown, uint32_t* flagUp, uint32_t* flagDown,
                            uint32_t* counter, uint32_t value, size_t planeBytes, cudaStream_t st);
cudaError_t launchPlaneCopy(const void* src, void* dst, const PlaneList& pl, size_t planeBytes, cudaStream_t st);

}  // namespace nlbm
