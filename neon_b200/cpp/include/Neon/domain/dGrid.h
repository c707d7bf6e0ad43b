// dGrid.h — dense grid, z-slab partitioned over the Backend's devices, structure-of-arrays fields with a 512-byte
// aligned row pitch, host mirror + device partitions, halo update.
//
// Mirrors libNeonDomain/include/Neon/domain/details/dGrid/:
//   partitioning  dGrid_imp.h:32-63  (floor(Z/n) planes each, the first Z mod n devices one more; x and y whole)
//   halo radius   dGrid_imp.h:65-71, dField_imp.h:48-51 (one ghost plane per side once there is more than one device)
//   field layout  dField_imp.h:67-87 (reference: unpadded SoA; here pop[q][zm][y][x] with the pitch of nlbm_dense_layout)
//   host side     FieldBase::forEachActiveCell / updateDeviceData / updateHostData / ioToVtk (interface/FieldBase_imp.h:97-141,299-319)
//   halo update   dField::newHaloUpdate (dField.h:84-87; dField_imp.h:341-421,548-641)
// Every device operation is a C-ABI call (include/neon_lbm.h) or a cudaMemcpy; nothing here computes.
#pragma once

#include <algorithm>
#include <cstring>
#include <fstream>
#include <functional>
#include <type_traits>

#include "Neon/Neon.h"
#include "Neon/set/Backend.h"
#include "Neon/set/Container.h"

namespace Neon::domain {

/* Neon::domain::Stencil (interface/Stencil.h:16-17): the list of neighbour offsets a grid must serve. */
class Stencil
{
   public:
    Stencil() = default;
    Stencil(const std::vector<index_3d>& points, bool filterCenterOut = false)
    {
        for (const auto& p : points) {
            if (!(filterCenterOut && p == index_3d(0, 0, 0))) {
                mPoints.push_back(p);
            }
        }
    }
    const std::vector<index_3d>& points() const { return mPoints; }
    int                          nPoints() const { return int(mPoints.size()); }
    int                          getRadius() const
    {
        int r = 0;
        for (const auto& p : mPoints) {
            r = std::max({r, std::abs(p.x), std::abs(p.y), std::abs(p.z)});
        }
        return r;
    }

   private:
    std::vector<index_3d> mPoints;
};

/* A field type whose device representation is the packed 32-bit flag word of the kernel library specialises this
 * (CellType does, Neon/lbm/CellType.h): static uint32_t pack(const T&), static T unpack(uint32_t). */
template <typename T, typename = void>
struct FlagWordCodec
{
    static constexpr bool enabled = false;
};

}  // namespace Neon::domain

namespace Neon::domain {
/* Neon::domain::NghData (interface/NghData.h:9-49): value of a neighbour cell + whether that neighbour exists */
template <typename T>
struct NghData
{
    T    mData;
    bool mIsValid;
    NEON_CUDA_HOST_DEVICE T    operator()() const { return mData; }
    NEON_CUDA_HOST_DEVICE T    getData() const { return mData; }
    NEON_CUDA_HOST_DEVICE bool isValid() const { return mIsValid; }
};
}  // namespace Neon::domain

namespace Neon {

template <typename T, int C>
class dField;

/* cell handle inside a partition: x, y and the LOCAL z (dIndex.h) */
struct dIdx
{
    int32_t x = 0, y = 0, z = 0;
};

/* which cells of a partition one launch visits (dSpan.h:45-48, dSpan_imp.h:6-43): STANDARD all local planes, INTERNAL
 * [r, nz-r), BOUNDARY [0,r) U [nz-r,nz) with r = z_halo (the reference's BOUNDARY fold is not reproduced) */
struct dSpan
{
    using Idx = dIdx;
    int32_t nx = 0, ny = 0, nzLocal = 0, r = 0, nzView = 0;
    int32_t view = 0;
    NEON_CUDA_HOST_DEVICE bool setAndValidate(dIdx& idx, int x, int y, int z) const
    {
        if (x >= nx || y >= ny || z >= nzView) {
            return false;
        }
        idx.x = x;
        idx.y = y;
        idx.z = view == int(DataView::STANDARD) ? z : view == int(DataView::INTERNAL) ? z + r : (z < r ? z : nzLocal - 2 * r + z);
        return true;
    }
};

namespace detail {

struct dGridState
{
    Backend          backend;
    index_3d         dim;
    domain::Stencil  stencil;
    int              nParts = 1;
    int              zHalo = 0;
    std::vector<int> nzLocal, zOrigin;
    int64_t          pitchY = 0; /* elements, common to every field of the grid */
    uint64_t         nextUid = 1;
};

inline bool sameOffsets(const std::vector<index_3d>& a, const int (*b)[3], int n)
{
    if (int(a.size()) != n) {
        return false;
    }
    for (int k = 0; k < n; ++k) {
        if (a[k].x != b[k][0] || a[k].y != b[k][1] || a[k].z != b[k][2]) {
            return false;
        }
    }
    return true;
}

/* 19 / 27 when the offsets are D3Q19 (benchmarks/lbm-lid-driven-cavity-flow/src/D3Q19.h:23-44) or D3Q27
 * (apps/lbmMultiRes/lattice.h:15-77: x slowest, then y, then z, each in the order 0, -1, +1) in the library's order */
inline int latticeOf(const std::vector<index_3d>& points)
{
    static const int d3q19[19][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1},
                                     {0, -1, -1}, {0, -1, 1}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0},
                                     {1, -1, 0}, {1, 0, 1}, {1, 0, -1}, {0, 1, 1}, {0, 1, -1}};
    if (sameOffsets(points, d3q19, 19)) {
        return 19;
    }
    int       d3q27[27][3];
    const int order[3] = {0, -1, 1};
    for (int k = 0; k < 27; ++k) {
        d3q27[k][0] = order[k / 9];
        d3q27[k][1] = order[(k / 3) % 3];
        d3q27[k][2] = order[k % 3];
    }
    return sameOffsets(points, d3q27, 27) ? 27 : 0;
}

}  // namespace detail

class dGrid
{
   public:
    template <typename T, int C = 0>
    using Field = dField<T, C>;

    dGrid() = default;

    /* Grid(bk, dim, activeLambda, stencil) — a dense grid keeps every cell; the lambda is accepted for API parity
     * (dGrid.h: "implicit" constructor) and must not deactivate cells. */
    template <typename ActiveCellLambda>
    dGrid(const Backend& bk, const index_3d& dim, ActiveCellLambda activeCellLambda, const domain::Stencil& stencil)
        : mS(std::make_shared<detail::dGridState>())
    {
        (void)activeCellLambda;
        auto& s = *mS;
        s.backend = bk;
        s.dim = dim;
        s.stencil = stencil;
        s.nParts = bk.getDeviceCount();
        if (dim.x <= 0 || dim.y <= 0 || dim.z <= 0 || dim.z < s.nParts) {
            NeonException e("dGrid");
            e << "cannot split a " << dim.to_string() << " box over " << s.nParts << " device(s)";
            NEON_THROW(e);
        }
        if (stencil.getRadius() > 1) {
            NEON_THROW_UNSUPPORTED_OPERATION("the LBM path serves radius-1 stencils (D3Q19 / D3Q27)");
        }
        s.zHalo = s.nParts > 1 ? 1 : 0;
        const int base = dim.z / s.nParts, rem = dim.z % s.nParts;
        int       origin = 0;
        for (int i = 0; i < s.nParts; ++i) {
            const int nz = base + (i < rem ? 1 : 0);
            s.nzLocal.push_back(nz);
            s.zOrigin.push_back(origin);
            origin += nz;
        }
        /* one row pitch (in elements) for every field of the grid, so that flag words and populations index alike:
         * the 4-byte pitch of the library (512-byte rows) — 1 KB rows for 8-byte fields */
        nlbm_dense_desc d = descOf(0);
        detail::check(nlbm_dense_layout(&d, 1, 4, nullptr, nullptr), "nlbm_dense_layout");
        s.pitchY = d.pitch_y;
    }

    const Backend&         getBackend() const { return mS->backend; }
    const index_3d&        getDimension() const { return mS->dim; }
    const domain::Stencil& getStencil() const { return mS->stencil; }
    int                    getNumPartitions() const { return mS->nParts; }
    size_t                 getNumActiveCells() const { return mS->dim.rMul<size_t>(); }
    int                    zHalo() const { return mS->zHalo; }
    int                    nzLocal(int setIdx) const { return mS->nzLocal.at(setIdx); }
    int                    zOrigin(int setIdx) const { return mS->zOrigin.at(setIdx); }
    bool                   isInsideDomain(const index_3d& p) const
    {
        return p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < mS->dim.x && p.y < mS->dim.y && p.z < mS->dim.z;
    }

    /* 19 or 27 if the grid's stencil is one of the two lattices of the kernel library in ITS order, else 0 */
    int latticeQ() const { return detail::latticeOf(mS->stencil.points()); }

    /* partition descriptor without pointers */
    nlbm_dense_desc descOf(int setIdx) const
    {
        const auto&     s = *mS;
        nlbm_dense_desc d{};
        d.nx = s.dim.x;
        d.ny = s.dim.y;
        d.nz_local = s.nzLocal.at(setIdx);
        d.z_halo = s.zHalo;
        d.pitch_y = s.pitchY;
        d.pitch_z = s.pitchY * s.dim.y;
        d.pitch_q = d.pitch_z * (d.nz_local + 2 * d.z_halo);
        d.z_origin = s.zOrigin.at(setIdx);
        d.gnx = s.dim.x;
        d.gny = s.dim.y;
        d.gnz = s.dim.z;
        return d;
    }

    template <typename T, int C = 0>
    dField<T, C> newField(const std::string& name, int cardinality, T outsideValue = T()) const
    {
        return dField<T, C>(*this, name, cardinality, outsideValue, mS->nextUid++);
    }

    using Span = dSpan;
    using Idx = dIdx;
    dSpan getSpan(int setIdx, DataView dataView) const
    {
        dSpan sp;
        sp.nx = mS->dim.x;
        sp.ny = mS->dim.y;
        sp.nzLocal = mS->nzLocal.at(setIdx);
        sp.r = mS->zHalo;
        sp.view = int(dataView);
        const int internal = std::max(0, sp.nzLocal - 2 * sp.r);
        sp.nzView = dataView == DataView::STANDARD ? sp.nzLocal : dataView == DataView::INTERNAL ? internal : sp.nzLocal - internal;
        return sp;
    }

    /* Grid::newContainer(name, loadingLambda): a container whose body is a per-cell device lambda, launched by a
     * generic kernel over the span of the data view (DeviceContainer.h:88-111, LambdaExecutor.h:12-39).  Needs nvcc with
     * --extended-lambda: include "Neon/domain/GenericContainer.h" from a .cu translation unit. */
    template <typename LoadingLambda>
    set::Container newContainer(const std::string& name, LoadingLambda loadingLambda) const;

   private:
    std::shared_ptr<detail::dGridState> mS;
};

namespace detail {

/* Halo update of one dense field: per device and face one nlbm_dense_halo_push, ordered by events only.
 *   ready[d]  : everything device d enqueued so far on the halo's stream (its boundary planes are final, and its earlier
 *               reads of the planes about to be overwritten are done)
 *   get       : device d waits for its neighbours' ready events and pulls their boundary planes into its ghost planes
 *   put       : device d waits likewise, pushes its boundary planes into the neighbours' ghost planes (NVLink stores),
 *               then everybody waits for the neighbours' done events
 * The reference fences the same copies with host-blocking stream syncs (SynchronizationContainer.h:37-42). */
struct DenseHaloImpl : set::Container::Impl
{
    dGrid                        grid;
    std::vector<void*>           mem;  /* device base of the field, per partition */
    int                          elemBytes = 4, cardinality = 1, latticeQ = 0;
    bool                         isFlagWords = false;
    set::TransferMode            mode = set::TransferMode::get;
    std::vector<cudaEvent_t>     ready, done;
    std::vector<nlbm_dense_desc> desc;

    void run(int streamIdx, DataView) override
    {
        const int n = grid.getNumPartitions();
        if (n == 1) {
            return;
        }
        if (backend.runtime() != Runtime::stream) {
            NeonException e(name);
            e << "halo updates need Runtime::stream";
            NEON_THROW(e);
        }
        if (ready.empty()) {
            for (int d = 0; d < n; ++d) {
                ready.push_back(backend.newEvent(d));
                done.push_back(backend.newEvent(d));
            }
        }
        for (int d = 0; d < n; ++d) {
            backend.setDevice(d);
            NEON_CUDA_CHECK(cudaEventRecord(ready[d], backend.stream(d, streamIdx)));
        }
        for (int d = 0; d < n; ++d) {
            backend.setDevice(d);
            cudaStream_t st = backend.stream(d, streamIdx);
            for (int nbr : {d - 1, d + 1}) {
                if (nbr < 0 || nbr >= n) {
                    continue; /* no periodic wrap, dField_imp.h:409-415 */
                }
                NEON_CUDA_CHECK(cudaStreamWaitEvent(st, ready[nbr], 0));
            }
            for (int nbr : {d - 1, d + 1}) {
                if (nbr < 0 || nbr >= n) {
                    continue;
                }
                if (mode == set::TransferMode::get) {
                    /* the neighbour below hands me its top plane (dir +1), the one above its bottom plane (dir -1) */
                    const int dir = nbr < d ? +1 : -1;
                    check(nlbm_dense_halo_push(&desc[nbr], mem[nbr], &desc[d], mem[d], elemBytes, cardinality, latticeQ, dir, st),
                          "nlbm_dense_halo_push");
                } else {
                    const int dir = nbr > d ? +1 : -1;
                    check(nlbm_dense_halo_push(&desc[d], mem[d], &desc[nbr], mem[nbr], elemBytes, cardinality, latticeQ, dir, st),
                          "nlbm_dense_halo_push");
                }
            }
        }
        if (mode == set::TransferMode::put) {
            for (int d = 0; d < n; ++d) {
                backend.setDevice(d);
                NEON_CUDA_CHECK(cudaEventRecord(done[d], backend.stream(d, streamIdx)));
            }
            for (int d = 0; d < n; ++d) {
                backend.setDevice(d);
                for (int nbr : {d - 1, d + 1}) {
                    if (nbr >= 0 && nbr < n) {
                        NEON_CUDA_CHECK(cudaStreamWaitEvent(backend.stream(d, streamIdx), done[nbr], 0));
                    }
                }
            }
        }
        if (isFlagWords) {
            for (int d = 0; d < n; ++d) {
                backend.setDevice(d);
                nlbm_dense_desc fd = desc[d];
                fd.flags = static_cast<uint32_t*>(mem[d]);
                check(nlbm_dense_flags_commit(&fd, backend.stream(d, streamIdx)), "nlbm_dense_flags_commit");
            }
        }
    }
    void run(int, int, DataView) override
    {
        NEON_THROW_UNSUPPORTED_OPERATION("a halo update involves every device; run it without a SetIdx");
    }
    size_t bytesPerFace() const
    {
        const int crossing = latticeQ == 19 ? 5 : latticeQ == 27 ? 9 : cardinality;
        return size_t(crossing) * size_t(desc[0].pitch_z) * size_t(elemBytes);
    }
};

}  // namespace detail

template <typename T, int C = 0>
class dField
{
    using Codec = domain::FlagWordCodec<T>;
    static constexpr bool kFlagWords = Codec::enabled;
    static_assert(kFlagWords || std::is_same_v<T, float> || std::is_same_v<T, double> || std::is_same_v<T, int32_t> ||
                      std::is_same_v<T, uint32_t>,
                  "dField: float, double, 32-bit integers, or a type with a FlagWordCodec");

   public:
    using Type = T;
    using Grid = dGrid;
    using DeviceType = std::conditional_t<kFlagWords, uint32_t, T>;

    /* What Loader::load returns: the device memory of the field on one device and its layout — the reference's
     * dPartition (dPartition.h:423-434).  The LBM containers hand `mem()` and `desc` to the C ABI; generic per-cell
     * lambdas (dGrid::newContainer, compiled by nvcc) use the accessors: operator()(idx, card), getNghData<dx,dy,dz>
     * (dPartition.h:174-207, 265-299), getGlobalIndex.  A flag-word field exposes the packed 32-bit words. */
    struct Partition
    {
        using Idx = dIdx;
        using Type = DeviceType;
        DeviceType*     memory = nullptr;
        nlbm_dense_desc desc{};
        int             card = 0;
        DeviceType      outsideValue{};

        NEON_CUDA_HOST_DEVICE DeviceType* mem() const { return memory; }
        NEON_CUDA_HOST_DEVICE index_3d    dim() const { return {desc.nx, desc.ny, desc.nz_local}; }
        NEON_CUDA_HOST_DEVICE index_3d    origin() const { return {0, 0, desc.z_origin}; }
        NEON_CUDA_HOST_DEVICE int         zHalo() const { return desc.z_halo; }
        NEON_CUDA_HOST_DEVICE int         cardinality() const { return card; }

        NEON_CUDA_HOST_DEVICE int64_t offset(int x, int y, int zLocal, int c) const
        {
            return int64_t(c) * desc.pitch_q + int64_t(zLocal + desc.z_halo) * desc.pitch_z + int64_t(y) * desc.pitch_y + x;
        }
        NEON_CUDA_HOST_DEVICE DeviceType& operator()(const dIdx& i, int c) const { return memory[offset(i.x, i.y, i.z, c)]; }
        NEON_CUDA_HOST_DEVICE index_3d    getGlobalIndex(const dIdx& i) const { return {i.x, i.y, i.z + desc.z_origin}; }
        /* a neighbour is valid when it lies inside the global box (its plane is then local or a ghost plane) */
        NEON_CUDA_HOST_DEVICE bool isNghValid(const dIdx& i, int dx, int dy, int dz) const
        {
            const int x = i.x + dx, y = i.y + dy, gz = i.z + dz + desc.z_origin, lz = i.z + dz;
            return x >= 0 && x < desc.nx && y >= 0 && y < desc.ny && gz >= 0 && gz < desc.gnz && lz >= -desc.z_halo &&
                   lz < desc.nz_local + desc.z_halo;
        }
        NEON_CUDA_HOST_DEVICE domain::NghData<DeviceType> getNghData(const dIdx& i, const index_3d& off, int c) const
        {
            domain::NghData<DeviceType> r;
            r.mIsValid = isNghValid(i, off.x, off.y, off.z);
            if (r.mIsValid) {
                r.mData = memory[offset(i.x + off.x, i.y + off.y, i.z + off.z, c)];
            } else {
                r.mData = outsideValue;
            }
            return r;
        }
        template <int dx, int dy, int dz>
        NEON_CUDA_HOST_DEVICE domain::NghData<DeviceType> getNghData(const dIdx& i, int c) const
        {
            return getNghData(i, index_3d(dx, dy, dz), c);
        }
        template <int dx, int dy, int dz>
        NEON_CUDA_HOST_DEVICE DeviceType getNghData(const dIdx& i, int c, DeviceType alternative) const
        {
            return isNghValid(i, dx, dy, dz) ? memory[offset(i.x + dx, i.y + dy, i.z + dz, c)] : alternative;
        }
    };
    using Idx = dIdx;

    dField() = default;

    const std::string& getName() const { return mS->name; }
    uint64_t           getUid() const { return mS->uid; }
    int                getCardinality() const { return mS->cardinality; }
    const dGrid&       getGrid() const { return mS->grid; }
    const index_3d&    getDimension() const { return mS->grid.getDimension(); }
    const Backend&     getBackend() const { return mS->grid.getBackend(); }
    Partition&         getPartition(int setIdx) { return mS->parts.at(setIdx); }
    const Partition&   getPartition(int setIdx) const { return mS->parts.at(setIdx); }
    bool isValid() const { return bool(mS); }

    // ---- host mirror ---------------------------------------------------------------------------------------------
    T& getReference(const index_3d& p, int card)
    {
        ensureHost();
        return mS->host[hostOffset(p, card)];
    }
    T operator()(const index_3d& p, int card) const
    {
        ensureHost();
        return mS->grid.isInsideDomain(p) ? mS->host[hostOffset(p, card)] : mS->outside;
    }
    T* hostData() /* [cardinality][z][y][x], unpadded */
    {
        ensureHost();
        return mS->host;
    }
    const T* hostData() const
    {
        ensureHost();
        return mS->host;
    }

    /* fn(const index_3d&, const int& cardinality, T&) over every cell of the host mirror */
    template <typename Fn>
    void forEachActiveCell(Fn fn, computeMode_t mode = computeMode_t::par)
    {
        ensureHost();
        const index_3d dim = getDimension();
        const int      card = mS->cardinality;
        T*             host = mS->host;
        const size_t   cells = dim.rMul<size_t>();
        if (mode == computeMode_t::par) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
            for (int z = 0; z < dim.z; ++z) {
                for (int y = 0; y < dim.y; ++y) {
                    for (int x = 0; x < dim.x; ++x) {
                        const index_3d p(x, y, z);
                        const size_t   o = (size_t(z) * dim.y + y) * dim.x + x;
                        for (int c = 0; c < card; ++c) {
                            fn(p, c, host[size_t(c) * cells + o]);
                        }
                    }
                }
            }
        } else {
            for (int z = 0; z < dim.z; ++z) {
                for (int y = 0; y < dim.y; ++y) {
                    for (int x = 0; x < dim.x; ++x) {
                        const index_3d p(x, y, z);
                        const size_t   o = (size_t(z) * dim.y + y) * dim.x + x;
                        for (int c = 0; c < card; ++c) {
                            fn(p, c, host[size_t(c) * cells + o]);
                        }
                    }
                }
            }
        }
    }
    template <typename Fn>
    void forEachActiveCell(Fn fn, computeMode_t mode = computeMode_t::par) const
    {
        const_cast<dField*>(this)->forEachActiveCell([&](const index_3d& p, const int& c, T& v) { fn(p, c, const_cast<const T&>(v)); },
                                                     mode);
    }

    // ---- x-face cache (include/neon_lbm.h): the LBM step keeps wall values of the cells at x = 0 / nx-1 from it.  Every
    // writer of this class refreshes it (updateDeviceData); whoever writes the device memory otherwise (set-up kernels,
    // generic containers) calls commitWalls() afterwards — generic containers invalidate it when they run.
    void commitWalls(int streamIdx = Backend::mainStreamIdx)
    {
        if constexpr (!kFlagWords) {
            const Backend& bk = getBackend();
            if (bk.runtime() != Runtime::stream) {
                return;
            }
            for (int d = 0; d < mS->grid.getNumPartitions(); ++d) {
                bk.setDevice(d);
                nlbm_dense_desc desc = mS->grid.descOf(d);
                desc.pop_out = mS->dev[d];
                desc.wall_cache = mS->wallCache[d];
                detail::check(nlbm_dense_wall_cache_build(&desc, mS->cardinality, int(sizeof(T)), bk.stream(d, streamIdx)),
                              "nlbm_dense_wall_cache_build");
            }
            mS->wallsCommitted = true;
        }
    }
    void  invalidateWalls() const { mS->wallsCommitted = false; }
    void* wallCachePtr(int setIdx) const { return mS->wallsCommitted ? mS->wallCache.at(setIdx) : nullptr; }

    // ---- host <-> device (FieldBase::updateDeviceData / updateHostData), asynchronous on stream `streamIdx` ----------
    void updateDeviceData(int streamIdx = Backend::mainStreamIdx)
    {
        transfer(streamIdx, /*toDevice*/ true);
        commitWalls(streamIdx);
    }
    void updateHostData(int streamIdx = Backend::mainStreamIdx)
    {
        transfer(streamIdx, /*toDevice*/ false);
    }

    // ---- dField::newHaloUpdate (dField.h:84-87) ---------------------------------------------------------------------
    set::Container newHaloUpdate(set::StencilSemantic semantic, set::TransferMode mode, Execution execution = Execution::device) const
    {
        if (execution != Execution::device) {
            NEON_THROW_UNSUPPORTED_OPERATION("host-side halo update: there is no CPU path");
        }
        auto         impl = std::make_shared<detail::DenseHaloImpl>();
        const dGrid& g = mS->grid;
        impl->grid = g;
        impl->backend = g.getBackend();
        impl->kind = set::Container::Kind::halo;
        impl->mode = mode;
        impl->elemBytes = int(sizeof(DeviceType));
        impl->cardinality = mS->cardinality;
        impl->isFlagWords = kFlagWords;
        impl->latticeQ = 0;
        if (semantic == set::StencilSemantic::streaming) {
            /* lattice semantic: component k travels along stencil point k; the library knows its two lattices
             * (the reference throws here with more than one device, dField_imp.h:610-612) */
            const int q = g.latticeQ();
            if (q == 0 || q != mS->cardinality) {
                NEON_THROW_UNSUPPORTED_OPERATION("streaming halo semantic needs a D3Q19/D3Q27 grid stencil and a field of that cardinality");
            }
            impl->latticeQ = q;
        }
        for (int d = 0; d < g.getNumPartitions(); ++d) {
            impl->mem.push_back(mS->dev[d]);
            impl->desc.push_back(g.descOf(d));
        }
        impl->name = "haloUpdate(" + mS->name + "," + set::StencilSemanticUtils::toString(semantic) + "," +
                     set::TransferModeUtils::toString(mode) + ")";
        set::Token t;
        t.uid = mS->uid;
        t.fieldName = mS->name;
        t.access = set::Access::write;
        impl->tokens.push_back(t);
        return set::Container(impl);
    }

    // ---- export (FieldBase::ioToVtk): legacy VTK, ASCII, voxel data, components interleaved ---------------------------
    void ioToVtk(const std::string& fileName, const std::string& fieldName, bool includeDomain = false,
                 IoFileType ioFileType = IoFileType::ASCII, bool isNodeSpace = false) const
    {
        (void)includeDomain;
        (void)isNodeSpace;
        if (ioFileType != IoFileType::ASCII) {
            NEON_THROW_UNSUPPORTED_OPERATION("binary VTK export");
        }
        if constexpr (kFlagWords) {
            NEON_THROW_UNSUPPORTED_OPERATION("VTK export of a flag field");
        } else {
            ensureHost();
            const index_3d dim = getDimension();
            std::ofstream  out(fileName + ".vtk", std::ios::out | std::ios::binary);
            if (!out) {
                NeonException e("ioToVtk");
                e << "cannot open " << fileName << ".vtk";
                NEON_THROW(e);
            }
            const size_t cells = dim.rMul<size_t>();
            out << "# vtk DataFile Version 3.0\nTitle Neon\nASCII\nDATASET STRUCTURED_POINTS\n";
            out << "DIMENSIONS " << dim.x + 1 << " " << dim.y + 1 << " " << dim.z + 1 << "\n";
            out << "ORIGIN 0 0 0\nSPACING 1 1 1\n";
            out << "CELL_DATA " << cells << "\n";
            out << "SCALARS " << fieldName << " " << (std::is_same_v<T, double> ? "double " : std::is_same_v<T, float> ? "float " : "int ");
            if (mS->cardinality != 1) {
                out << " " << mS->cardinality;
            }
            out << "\nLOOKUP_TABLE default\n";
            for (size_t o = 0; o < cells; ++o) {
                for (int c = 0; c < mS->cardinality; ++c) {
                    out << mS->host[size_t(c) * cells + o] << (c + 1 < mS->cardinality ? " " : "\n");
                }
            }
            out << "METADATA\nINFORMATION 0\n\n";
        }
    }

   private:
    friend class dGrid;
    struct State
    {
        dGrid                 grid;
        std::string           name;
        int                   cardinality = 0;
        T                     outside{};
        uint64_t              uid = 0;
        T*                    host = nullptr;
        bool                  hostPinned = false;
        std::vector<void*>    dev;      /* per partition */
        std::vector<Partition> parts;   /* per partition: what Loader::load hands out */
        std::vector<void*>    wallCache; /* per partition: x-face cache of the field (nlbm_dense_wall_cache_build) */
        bool                  wallsCommitted = false;
        std::vector<size_t>   devBytes; /* per partition */
        std::vector<uint32_t*> staging; /* flag-word fields: pinned [zm][y][pitch] words per partition */
        ~State()
        {
            const Backend& bk = grid.getBackend();
            for (size_t d = 0; d < dev.size(); ++d) {
                if (dev[d]) {
                    cudaSetDevice(bk.devId(int(d)));
                    cudaFree(dev[d]);
                }
            }
            for (size_t d = 0; d < wallCache.size(); ++d) {
                if (wallCache[d]) {
                    cudaSetDevice(bk.devId(int(d)));
                    cudaFree(wallCache[d]);
                }
            }
            for (auto* p : staging) {
                if (p) {
                    cudaFreeHost(p);
                }
            }
            if (host) {
                if (hostPinned) {
                    cudaFreeHost(host);
                } else {
                    delete[] host;
                }
            }
        }
    };

    dField(const dGrid& grid, const std::string& name, int cardinality, T outside, uint64_t uid) : mS(std::make_shared<State>())
    {
        if (C != 0 && cardinality != C) {
            NeonException e("dField");
            e << "cardinality " << cardinality << " does not match the static cardinality " << C;
            NEON_THROW(e);
        }
        if (cardinality < 1 || cardinality > 27 || (kFlagWords && cardinality != 1)) {
            NeonException e("dField");
            e << "unsupported cardinality " << cardinality;
            NEON_THROW(e);
        }
        auto& s = *mS;
        s.grid = grid;
        s.name = name;
        s.cardinality = cardinality;
        s.outside = outside;
        s.uid = uid;
        const Backend& bk = grid.getBackend();
        const bool     cuda = bk.runtime() == Runtime::stream;
        /* the host mirror is allocated at its first use (ensureHost): a run that sets the problem up on the device and
         * never reads fields back needs none */
        for (int d = 0; d < grid.getNumPartitions(); ++d) {
            nlbm_dense_desc desc = grid.descOf(d);
            size_t          bytes = 0;
            if constexpr (kFlagWords) {
                nlbm_dense_desc tmp = desc;
                size_t          popBytes = 0;
                detail::check(nlbm_dense_layout(&tmp, 1, 4, &popBytes, &bytes), "nlbm_dense_layout");
            } else {
                bytes = size_t(cardinality) * size_t(desc.pitch_q) * sizeof(T);
            }
            void* p = nullptr;
            if (cuda) {
                bk.setDevice(d);
                NEON_CUDA_CHECK(cudaMalloc(&p, bytes));
                NEON_CUDA_CHECK(cudaMemset(p, 0, bytes));
            }
            s.dev.push_back(p);
            s.devBytes.push_back(bytes);
            Partition part;
            part.memory = static_cast<DeviceType*>(p);
            part.desc = desc;
            part.card = cardinality;
            if constexpr (kFlagWords) {
                part.outsideValue = Codec::pack(outside);
            } else {
                part.outsideValue = outside;
            }
            s.parts.push_back(part);
            void* wc = nullptr;
            if constexpr (!kFlagWords) {
                if (cuda) {
                    size_t wcBytes = 0;
                    detail::check(nlbm_dense_wall_cache_layout(&desc, cardinality, int(sizeof(T)), &wcBytes), "nlbm_dense_wall_cache_layout");
                    NEON_CUDA_CHECK(cudaMalloc(&wc, wcBytes));
                }
            }
            s.wallCache.push_back(wc);
            if constexpr (kFlagWords) {
                uint32_t* st = nullptr;
                if (cuda) {
                    const size_t words = size_t(desc.pitch_z) * size_t(desc.nz_local + 2 * desc.z_halo);
                    NEON_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&st), words * 4, cudaHostAllocDefault));
                }
                s.staging.push_back(st);
            }
        }
    }

    /* pinned when it will be copied to a device (asynchronous, full-rate transfers); filled with the outside value */
    void ensureHost() const
    {
        auto& s = *mS;
        if (s.host) {
            return;
        }
        const size_t n = s.grid.getNumActiveCells() * size_t(s.cardinality);
        const bool   cuda = s.grid.getBackend().runtime() == Runtime::stream;
        if (cuda && cudaHostAlloc(reinterpret_cast<void**>(&s.host), n * sizeof(T), cudaHostAllocDefault) == cudaSuccess) {
            s.hostPinned = true;
        } else {
            (void)cudaGetLastError();
            s.host = new T[n];
        }
        const T outside = s.outside;
        T*      host = s.host;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int64_t i = 0; i < int64_t(n); ++i) {
            host[i] = outside;
        }
    }

    size_t hostOffset(const index_3d& p, int card) const
    {
        const index_3d& dim = mS->grid.getDimension();
        return (size_t(card) * dim.z + p.z) * size_t(dim.y) * dim.x + size_t(p.y) * dim.x + p.x;
    }

    void transfer(int streamIdx, bool toDevice)
    {
        const dGrid&   g = mS->grid;
        const Backend& bk = g.getBackend();
        ensureHost();
        if (bk.runtime() != Runtime::stream) {
            return; /* host-only backend: the mirror is the field */
        }
        const index_3d dim = g.getDimension();
        for (int d = 0; d < g.getNumPartitions(); ++d) {
            bk.setDevice(d);
            cudaStream_t          st = bk.stream(d, streamIdx);
            const nlbm_dense_desc desc = g.descOf(d);
            /* memory planes [zm0, zm0 + n) <-> global planes [gz0, gz0 + n): to the device the in-box ghost planes go too */
            int zm0 = toDevice ? 0 : desc.z_halo, n = toDevice ? desc.nz_local + 2 * desc.z_halo : desc.nz_local;
            int gz0 = desc.z_origin - desc.z_halo + zm0;
            if (gz0 < 0) {
                zm0 -= gz0;
                n += gz0;
                gz0 = 0;
            }
            n = std::min(n, dim.z - gz0);
            if constexpr (kFlagWords) {
                uint32_t*    words = mS->staging[d];
                const size_t planeWords = size_t(desc.pitch_z);
                const int    nzm = desc.nz_local + 2 * desc.z_halo;
                if (toDevice) {
                    std::fill(words, words + planeWords * nzm, uint32_t(NLBM_UNDEFINED) << NLBM_FLAG_CLASS_SHIFT);
                    for (int i = 0; i < n; ++i) {
                        for (int y = 0; y < dim.y; ++y) {
                            const T*  src = mS->host + (size_t(gz0 + i) * dim.y + y) * dim.x;
                            uint32_t* dst = words + size_t(zm0 + i) * planeWords + size_t(y) * desc.pitch_y;
                            for (int x = 0; x < dim.x; ++x) {
                                dst[x] = Codec::pack(src[x]);
                            }
                        }
                    }
                    NEON_CUDA_CHECK(cudaMemcpyAsync(mS->dev[d], words, planeWords * nzm * 4, cudaMemcpyHostToDevice, st));
                    nlbm_dense_desc fd = desc;
                    fd.flags = static_cast<uint32_t*>(mS->dev[d]);
                    detail::check(nlbm_dense_flags_commit(&fd, st), "nlbm_dense_flags_commit");
                } else {
                    NEON_CUDA_CHECK(cudaMemcpyAsync(words, mS->dev[d], planeWords * nzm * 4, cudaMemcpyDeviceToHost, st));
                    NEON_CUDA_CHECK(cudaStreamSynchronize(st)); /* the unpack below reads the staging buffer */
                    for (int i = 0; i < n; ++i) {
                        for (int y = 0; y < dim.y; ++y) {
                            T*              dst = mS->host + (size_t(gz0 + i) * dim.y + y) * dim.x;
                            const uint32_t* src = words + size_t(zm0 + i) * planeWords + size_t(y) * desc.pitch_y;
                            for (int x = 0; x < dim.x; ++x) {
                                dst[x] = Codec::unpack(src[x]);
                            }
                        }
                    }
                }
            } else {
                const size_t cells = dim.rMul<size_t>();
                for (int c = 0; c < mS->cardinality; ++c) {
                    T* hostPtr = mS->host + size_t(c) * cells + size_t(gz0) * dim.y * dim.x;
                    T* devPtr = static_cast<T*>(mS->dev[d]) + size_t(c) * desc.pitch_q + size_t(zm0) * desc.pitch_z;
                    /* rows of consecutive planes are pitch_y apart on the device (pitch_z = pitch_y * ny): one 2-D copy */
                    if (toDevice) {
                        NEON_CUDA_CHECK(cudaMemcpy2DAsync(devPtr, size_t(desc.pitch_y) * sizeof(T), hostPtr, size_t(dim.x) * sizeof(T),
                                                          size_t(dim.x) * sizeof(T), size_t(dim.y) * n, cudaMemcpyHostToDevice, st));
                    } else {
                        NEON_CUDA_CHECK(cudaMemcpy2DAsync(hostPtr, size_t(dim.x) * sizeof(T), devPtr, size_t(desc.pitch_y) * sizeof(T),
                                                          size_t(dim.x) * sizeof(T), size_t(dim.y) * n, cudaMemcpyDeviceToHost, st));
                    }
                }
            }
        }
    }

    std::shared_ptr<State> mS;
};

}  // namespace Neon
