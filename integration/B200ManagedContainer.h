// integration/B200ManagedContainer.h — a device-managed Neon container: its body, per device, is a host function that
// enqueues device work itself on the stream the scheduler chose.
//
// Part of the reference-side binding of libneon_lbm.so: the piece of code a Neon maintainer adds next to
// the benchmark.  Neon v0.3.3 declares this kind of container (Container::factoryDeviceManaged, libNeonSet/include/Neon/set/
// Containter.h:110-114 -> container/DeviceManagedContainer.h) but never instantiates it, and that header no longer compiles
// against the current ContainerAPI: it calls `this->m_loadingLambda` (the member is `mLoadingLambda`, :70 vs :120),
// `getContainerType()` (ContainerAPI.h:107 has `getContainerExecutionType()`), reads the private `mParsingDataUpdated`
// (ContainerAPI.h:203-213 offers isParsingDataUpdated / setParsingDataUpdated) and constructs `Loader` with a
// Neon::DeviceType where Loader.h:49 takes a Neon::Execution.  This header is the same idea written against the interfaces
// that DeviceContainer.h:48-79 (the stock container) uses, so the tokens it reports come from the same Loader::load calls
// and the Skeleton treats it like the container it replaces.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "Neon/set/container/ContainerAPI.h"
#include "Neon/set/container/Loader.h"

namespace nlbm_shim {

template <typename DataContainer, typename Launch /* void(int streamIdx, Neon::DataView) */>
struct ManagedContainer : Neon::set::internal::ContainerAPI
{
    using Loader = Neon::set::Loader;

    ManagedContainer(const std::string& name, const DataContainer& grid, std::function<Launch(Neon::SetIdx, Loader&)> loading)
        : mLoading(std::move(loading)), mGrid(grid)
    {
        setName(name);
        setContainerExecutionType(Neon::set::ContainerExecutionType::deviceManaged);
        setContainerOperationType(Neon::set::ContainerOperationType::compute);
        setDataViewSupport(Neon::set::internal::ContainerAPI::DataViewSupport::on);
        this->parse();
    }
    ~ManagedContainer() override = default;

    auto parse() -> const std::vector<Neon::set::dataDependency::Token>& override
    {
        if (!this->isParsingDataUpdated()) {
            Loader parser(*this, Neon::Execution::host, Neon::SetIdx(0), Neon::DataView::STANDARD,
                          Neon::set::internal::LoadingMode_e::PARSE_AND_EXTRACT_LAMBDA);
            mLoading(Neon::SetIdx(0), parser);
            this->setParsingDataUpdated(true);
            this->setContainerPattern(this->getTokens());
        }
        return getTokens();
    }

    auto run(int streamIdx = 0, Neon::DataView dataView = Neon::DataView::STANDARD) -> void override
    {
        const int n = mGrid.getBackend().devSet().setCardinality();
        for (int i = 0; i < n; ++i) {  // every call below only enqueues: no host thread per device needed
            run(Neon::SetIdx(i), streamIdx, dataView);
        }
    }

    auto run(Neon::SetIdx setIdx, int streamIdx = 0, Neon::DataView dataView = Neon::DataView::STANDARD) -> void override
    {
        Loader loader(*this, Neon::Execution::device, setIdx, dataView, Neon::set::internal::LoadingMode_e::EXTRACT_LAMBDA);
        Launch launch = mLoading(setIdx, loader);
        launch(streamIdx, dataView);
    }

   private:
    std::function<Launch(Neon::SetIdx, Loader&)> mLoading;
    DataContainer                                mGrid;
};

// Neon::set::Container can only be made from a ContainerAPI by its own static factories (its constructors from a
// shared_ptr<ContainerAPI> are protected, Containter.h:161-166); a derived type reaches them
struct ContainerFromApi : Neon::set::Container
{
    explicit ContainerFromApi(std::shared_ptr<Neon::set::internal::ContainerAPI>& p) : Neon::set::Container(p) {}
};

template <typename DataContainer, typename LoadingLambda>
auto newManagedContainer(const std::string& name, const DataContainer& grid, const LoadingLambda& loading) -> Neon::set::Container
{
    using Launch = typename std::invoke_result<LoadingLambda, Neon::SetIdx, Neon::set::Loader&>::type;
    std::shared_ptr<Neon::set::internal::ContainerAPI> p(new ManagedContainer<DataContainer, Launch>(name, grid, loading));
    ContainerFromApi c(p);
    return static_cast<const Neon::set::Container&>(c);
}

}  // namespace nlbm_shim
