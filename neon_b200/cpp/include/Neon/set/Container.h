// Container.h — a unit of work the Skeleton schedules: run(streamIdx, dataView), plus the data tokens it declared.
//
// Mirrors Neon::set::Container (libNeonSet/include/Neon/set/Containter.h:16-167): run(int streamIdx, DataView)
// (:25-27) is THE drop-in boundary of the LBM path; below it the reference goes DeviceContainer::run
// (container/DeviceContainer.h:88-111) -> DevSet::launchLambdaOnSpan (DevSet.h:226-261) -> cudaLaunchKernel, here the
// body enqueues C-ABI calls (include/neon_lbm.h).  The loading lambda / Loader::load(field, Pattern, StencilSemantic)
// protocol that lets the Skeleton discover data dependencies (container/Loader.h:67-82, Loader_imp.h:52-96) is kept:
// factoryDeviceManaged is the analogue of Container::factoryDeviceManaged (Containter.h:110-114).
#pragma once

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "Neon/Neon.h"
#include "Neon/set/Backend.h"

namespace Neon::set {

enum class StencilSemantic
{
    standard = 0, /* "grid": every component crosses the face (--huGrid) */
    streaming = 1 /* "lattice": only the populations whose c_z points across the face (--huLattice) */
};
struct StencilSemanticUtils
{
    static std::string toString(StencilSemantic s) { return s == StencilSemantic::streaming ? "streaming" : "grid"; }
};
enum class TransferMode
{
    put = 0,
    get = 1
};
struct TransferModeUtils
{
    static std::string toString(TransferMode m) { return m == TransferMode::get ? "get" : "put"; }
};
enum class Access
{
    read,
    write
};

class Container;

/* what Loader::load recorded about one field (the reference's internal::dependencyTools::DataToken) */
struct Token
{
    uint64_t        uid = 0;
    std::string     fieldName;
    Access          access = Access::read;
    Pattern         pattern = Pattern::MAP;
    StencilSemantic semantic = StencilSemantic::standard;
    /* builds the halo-update container of the field (FieldBase::newHaloUpdate) — only set for stencil reads */
    std::function<Container(StencilSemantic, TransferMode, Execution)> newHaloUpdate;
    /* set for writes: tells the field that a container with an arbitrary body is about to write it (derived data the
     * library keeps about the field, e.g. dField's x-face cache, is then stale) */
    std::function<void()> onGenericWrite;
};

class Loader;

class Container
{
   public:
    enum class Kind
    {
        compute,
        halo
    };
    struct Impl
    {
        std::string        name;
        Kind               kind = Kind::compute;
        std::vector<Token> tokens;
        Backend            backend;
        virtual ~Impl() = default;
        virtual void run(int streamIdx, DataView dataView) = 0;
        /* per-device variant (Container::run(SetIdx, streamIdx, dataView), DeviceContainer.h:118-150) */
        virtual void run(int setIdx, int streamIdx, DataView dataView) = 0;
        /* false if running the container changes host-side state that a replayed CUDA graph would not repeat */
        virtual bool graphSafe() const { return true; }
    };

    Container() = default;
    explicit Container(std::shared_ptr<Impl> impl) : mImpl(std::move(impl)) {}

    void run(int streamIdx = Backend::mainStreamIdx, DataView dataView = DataView::STANDARD) const
    {
        need();
        mImpl->run(streamIdx, dataView);
    }
    void run(SetIdx setIdx, int streamIdx, DataView dataView = DataView::STANDARD) const
    {
        need();
        mImpl->run(setIdx.idx, streamIdx, dataView);
    }
    bool                      graphSafe() const { return need()->graphSafe(); }
    const std::string&        getName() const { return need()->name; }
    Kind                      getKind() const { return need()->kind; }
    const std::vector<Token>& getTokens() const { return need()->tokens; }
    const Backend&            getBackend() const { return need()->backend; }
    bool                      isValid() const { return bool(mImpl); }
    template <typename T>
    std::shared_ptr<T> as() const
    {
        return std::dynamic_pointer_cast<T>(mImpl);
    }

    /* A container whose body, per device, is a host function that enqueues device work itself.  `loading` is called once
     * per device with a Loader (it declares the fields through Loader::load and returns the launcher). */
    template <typename LoadingLambda>
    static Container factoryDeviceManaged(const std::string& name, const Backend& bk, LoadingLambda loading);

   private:
    Impl* need() const
    {
        if (!mImpl) {
            NeonException e("Container");
            e << "empty container";
            NEON_THROW(e);
        }
        return mImpl.get();
    }
    std::shared_ptr<Impl> mImpl;
};

/* Loader (libNeonSet/include/Neon/set/container/Loader.h:67-82): load(field [, Pattern::STENCIL, semantic]) returns the
 * partition of `field` on the device being loaded and records a token.  A const field is a read, a non-const one a
 * write (Loader_imp.h:52-187). */
class Loader
{
   public:
    Loader(int setIdx, std::vector<Token>* tokens) : mSetIdx(setIdx), mTokens(tokens) {}

    template <typename Field>
    auto load(Field& field, Pattern pattern = Pattern::MAP, StencilSemantic semantic = StencilSemantic::standard)
        -> std::conditional_t<std::is_const_v<Field>, const typename std::remove_const_t<Field>::Partition&,
                              typename std::remove_const_t<Field>::Partition&>
    {
        constexpr bool isConst = std::is_const_v<Field>;
        if (mTokens && mSetIdx == 0) {
            Token t;
            t.uid = field.getUid();
            t.fieldName = field.getName();
            t.access = isConst ? Access::read : Access::write;
            t.pattern = pattern;
            t.semantic = semantic;
            if (pattern == Pattern::STENCIL) {
                auto copy = field; /* fields are shallow handles (dField.h:150-195) */
                t.newHaloUpdate = [copy](StencilSemantic s, TransferMode m, Execution e) { return copy.newHaloUpdate(s, m, e); };
            }
            if (!isConst) {
                auto copy = field;
                t.onGenericWrite = [copy]() { copy.invalidateWalls(); };
            }
            mTokens->push_back(std::move(t));
        }
        return field.getPartition(mSetIdx);
    }
    int setIdx() const { return mSetIdx; }

   private:
    int                 mSetIdx;
    std::vector<Token>* mTokens;
};

namespace detail {
struct DeviceManagedImpl : Container::Impl
{
    std::vector<std::function<void(int, DataView)>> launchers; /* one per device */
    bool                                            genericBody = false; /* user lambda: may write anything it loaded non-const */
    /* optional: called once after every device has been issued (all-device run only) — e.g. to read back and check
     * per-device error counters without serialising the devices behind host syncs */
    std::function<void(int)> afterAll;
    void run(int streamIdx, DataView dataView) override
    {
        for (int d = 0; d < int(launchers.size()); ++d) {
            run(d, streamIdx, dataView);
        }
        if (afterAll) {
            afterAll(streamIdx);
        }
    }
    /* both entry points pass here: a user lambda that writes a population field makes that field's x-face cache stale
     * whichever way it was launched (DeviceContainer.h:118-150 lets callers run one device at a time) */
    void run(int setIdx, int streamIdx, DataView dataView) override
    {
        if (backend.runtime() != Runtime::stream) {
            NeonException e(name);
            e << "compute containers need Runtime::stream: there is no CPU fallback behind this veneer";
            NEON_THROW(e);
        }
        if (genericBody) {
            for (const auto& t : tokens) {
                if (t.access == Access::write && t.onGenericWrite) {
                    t.onGenericWrite();
                }
            }
        }
        backend.setDevice(setIdx);
        launchers.at(setIdx)(streamIdx, dataView);
    }
    /* true if running this container changes host-side state that a replayed CUDA graph would not repeat */
    bool graphSafe() const override { return !writesCachedFields(); }
    bool writesCachedFields() const
    {
        if (!genericBody) {
            return false;
        }
        for (const auto& t : tokens) {
            if (t.access == Access::write && t.onGenericWrite) {
                return true;
            }
        }
        return false;
    }
};
}  // namespace detail

template <typename LoadingLambda>
Container Container::factoryDeviceManaged(const std::string& name, const Backend& bk, LoadingLambda loading)
{
    auto impl = std::make_shared<detail::DeviceManagedImpl>();
    impl->name = name;
    impl->backend = bk;
    for (int d = 0; d < bk.getDeviceCount(); ++d) {
        Loader loader(d, &impl->tokens);
        impl->launchers.emplace_back(loading(SetIdx(d), loader));
    }
    return Container(impl);
}

}  // namespace Neon::set
