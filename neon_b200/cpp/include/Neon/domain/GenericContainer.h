// GenericContainer.h — Grid::newContainer(name, loadingLambda): containers whose body is an arbitrary per-cell device
// lambda, on the padded SoA layout of this library.
//
// Mirrors the reference's generic launch path: Grid::newContainer -> DeviceContainer<Grid,Lambda>::run
// (libNeonSet/include/Neon/set/container/DeviceContainer.h:88-111) -> DevSet::launchLambdaOnSpan (DevSet.h:226-261)
// -> the generic __global__ launchLambdaOnSpanCUDA (LambdaExecutor.h:12-39) over a dSpan (dSpan_imp.h:6-43).
// Differences by design: the loading lambda is parsed ONCE per device when the container is built (the reference
// re-runs it at every launch) and the launch goes straight to cudaLaunch (the reference calls cudaFuncGetAttributes on
// every launch, libNeonSys/include/Neon/sys/devices/gpu/GpuDevice.h:151-191), so a Skeleton of such containers can be
// captured into a CUDA graph.  bGrid::newContainer does the same over the blocks of a view (LambdaExecutor.h:105-116,
// bSpan_imp.h:7-21, bPartition_imp.h:97-124,194-198,340-358).  This header needs nvcc (--extended-lambda); the LBM hot path
// does not use it.
#pragma once

#ifndef __CUDACC__
#error "Neon/domain/GenericContainer.h defines device kernels: include it from a .cu translation unit (nvcc --extended-lambda)"
#endif

#include "Neon/domain/bGrid.h"
#include "Neon/domain/dGrid.h"

namespace Neon::detail {

template <typename Span, typename UserLambda>
__global__ void neonLambdaOnSpan(const Span span, UserLambda userLambda)
{
    typename Span::Idx idx;
    if (span.setAndValidate(idx, int(blockIdx.x * blockDim.x + threadIdx.x), int(blockIdx.y * blockDim.y + threadIdx.y),
                            int(blockIdx.z * blockDim.z + threadIdx.z))) {
        userLambda(idx);
    }
}

/* holds the parsed user lambda of one device; a plain struct so that the closure type never has to live in a host lambda */
template <typename Span, typename UserLambda>
struct SpanLauncher
{
    UserLambda fn;
    explicit SpanLauncher(const UserLambda& f) : fn(f) {}
    void launch(const Span& span, cudaStream_t stream) const
    {
        if (span.nzView <= 0) {
            return;
        }
        /* x-runs of one warp or more: 256 x 1 threads on long rows, 32 x 8 on short ones (reference: 256 x 1 x 1, dGrid_imp.h:16) */
        const unsigned bx = span.nx >= 256 ? 256u : span.nx > 64 ? 128u : span.nx > 32 ? 64u : 32u;
        const dim3     block(bx, 256u / bx, 1);
        const dim3     grid((span.nx + block.x - 1) / block.x, (span.ny + block.y - 1) / block.y, unsigned(span.nzView));
        neonLambdaOnSpan<Span, UserLambda><<<grid, block, 0, stream>>>(span, fn);
        NEON_CUDA_CHECK(cudaGetLastError());
    }
};

/* block-sparse grids: one CUDA block of 8 x 8 x 8 threads per block of the view (the reference launches 512-thread blocks
 * the same way, bSpan_imp.h:7-21) */
template <typename UserLambda>
__global__ void __launch_bounds__(512) neonLambdaOnBlocks(const bSpan span, UserLambda userLambda)
{
    bIdx idx;
    if (span.setAndValidate(idx, blockIdx.x, int(threadIdx.x), int(threadIdx.y), int(threadIdx.z))) {
        userLambda(idx);
    }
}
template <typename UserLambda>
struct BlockLauncher
{
    UserLambda fn;
    explicit BlockLauncher(const UserLambda& f) : fn(f) {}
    void launch(const bSpan& span, cudaStream_t stream) const
    {
        if (span.nBlocksView() == 0) {
            return;
        }
        neonLambdaOnBlocks<UserLambda><<<span.nBlocksView(), dim3(8, 8, 8), 0, stream>>>(span, fn);
        NEON_CUDA_CHECK(cudaGetLastError());
    }
};

}  // namespace Neon::detail

namespace Neon {

template <typename LoadingLambda>
set::Container bGrid::newContainer(const std::string& name, LoadingLambda loadingLambda) const
{
    auto impl = std::make_shared<set::detail::DeviceManagedImpl>();
    impl->name = name;
    impl->backend = getBackend();
    impl->genericBody = true;
    const Backend bk = getBackend();
    const bGrid   grid = *this;
    for (int d = 0; d < bk.getDeviceCount(); ++d) {
        set::Loader loader(d, &impl->tokens);
        auto        userLambda = loadingLambda(loader);
        using Launcher = detail::BlockLauncher<decltype(userLambda)>;
        auto launcher = std::make_shared<Launcher>(userLambda);
        impl->launchers.emplace_back([launcher, grid, bk, d](int streamIdx, DataView dataView) {
            launcher->launch(grid.getSpan(d, dataView), bk.stream(d, streamIdx));
        });
    }
    return set::Container(impl);
}

template <typename LoadingLambda>
set::Container dGrid::newContainer(const std::string& name, LoadingLambda loadingLambda) const
{
    auto impl = std::make_shared<set::detail::DeviceManagedImpl>();
    impl->name = name;
    impl->backend = getBackend();
    impl->genericBody = true;
    const Backend bk = getBackend();
    const dGrid   grid = *this;
    for (int d = 0; d < bk.getDeviceCount(); ++d) {
        set::Loader loader(d, &impl->tokens);
        auto        userLambda = loadingLambda(loader);
        using Launcher = detail::SpanLauncher<dSpan, decltype(userLambda)>;
        auto launcher = std::make_shared<Launcher>(userLambda);
        impl->launchers.emplace_back([launcher, grid, bk, d](int streamIdx, DataView dataView) {
            launcher->launch(grid.getSpan(d, dataView), bk.stream(d, streamIdx));
        });
    }
    return set::Container(impl);
}

}  // namespace Neon
