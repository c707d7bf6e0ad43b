// lbm_halo.cu — halo update of dense z-slab partitions.
//
// Replaces dField::newHaloUpdate / initHaloUpdateTable (libNeonDomain/.../dGrid/dField_imp.h:341-421,548-641) executed by
// DataTransferContainer::run (libNeonSet/include/Neon/set/container/DataTransferContainer.h:38-55): the reference issues
// one cudaMemcpyPeerAsync per population per direction behind a host-blocking stream sync.  Here ONE launch moves the
// planes of the populations that actually cross the face (c_z == dir).  Every population plane of a z-slice is a
// contiguous run of pitch_z elements, so the copy is a pure 16-byte streaming copy; with dst mapped from a peer GPU the
// stores travel over NVLink.
#include "lbm_common.cuh"
#include "lbm_host.h"

namespace nlbm {

__global__ void __launch_bounds__(256) k_plane_copy(const char* __restrict__ src, char* __restrict__ dst, const PlaneList pl,
                                                    const size_t vecPerPlane)
{
    const int    p = blockIdx.y;
    const uint4* s = reinterpret_cast<const uint4*>(src + pl.src[p]);
    uint4*       d = reinterpret_cast<uint4*>(dst + pl.dst[p]);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vecPerPlane; i += stride)
        d[i] = __ldcs(s + i);
}

cudaError_t launchPlaneCopy(const void* src, void* dst, const PlaneList& pl, size_t planeBytes, cudaStream_t st)
{
    if (pl.n == 0 || planeBytes == 0)
        return cudaSuccess;
    const size_t vecs = planeBytes / 16;
    size_t       bx = (vecs + 256 * 4 - 1) / (256 * 4);
    if (bx > 148 * 4)
        bx = 148 * 4;
    if (bx == 0)
        bx = 1;
    dim3 grid((unsigned)bx, pl.n);
    k_plane_copy<<<grid, 256, 0, st>>>((const char*)src, (char*)dst, pl, vecs);
    return cudaGetLastError();
}

}  // namespace nlbm
