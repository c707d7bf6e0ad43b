// bGrid.h — block-sparse grid of 8 x 8 x 8-cell blocks, z-partitioned by block layers over the Backend's devices.
//
// Mirrors libNeonDomain/include/Neon/domain/details/bGrid/ (Neon::bGrid = StaticBlock<8,8,8>, domain/bGrid.h:5):
//   blocks        bGrid_imp.h:7-185     a block exists when any of its cells is active (activeCellLambda)
//   connectivity  bGrid_imp.h:140-185, bPartition_imp.h:194-198   27 neighbour ids per block, index (dx+1)+3(dy+1)+9(dz+1)
//   active mask   StaticBlock.h:47-103  (on the device folded into the flag word: inactive cells carry class UNDEFINED)
//   partitioning  tools/partitioning/SpanDecomposition.h:98-150: 1-D over z by block layers, boundary blocks = the first and
//                 last layer of a partition; here the layers are split evenly (floor/ceil), at least two per device
//   halo update   bField_imp.h:173-332 (upstream ignores the cardinality — NaN for Q = 19); here only the
//                 facing z-slice of the crossing populations of every boundary block moves (nlbm_block_halo_push)
// Device layout: pop[q][blk][z][y][x], flags[blk][z][y][x], info[blk][32] (include/neon_lbm.h, nlbm_block_desc).
// The host mirror of a field is the dense box [cardinality][z][y][x] (inactive cells hold the outside value).
#pragma once

#include <algorithm>
#include <array>
#include <cstring>

#include "Neon/Neon.h"
#include "Neon/domain/dGrid.h"
#include "Neon/set/Backend.h"
#include "Neon/set/Container.h"

namespace Neon {

template <typename T, int C>
class bField;

namespace detail {

constexpr int kB = 8, kBlockCells = 512;

struct bPartitionInfo
{
    uint32_t                        nBlocks = 0, nAlloc = 0, nDown = 0, nUp = 0, nGhostDown = 0, nGhostUp = 0;
    std::vector<std::array<int, 3>> coords; /* (bz, by, bx) of local blocks, then ghost-down, then ghost-up blocks */
    uint32_t*                       infoDev = nullptr;
    uint32_t*                       activeMaskDev = nullptr;
};

struct bGridState
{
    Backend                     backend;
    index_3d                    dim, nb;
    domain::Stencil             stencil;
    std::vector<bPartitionInfo> parts;
    bool                        allActive = true;
    std::vector<uint8_t>        cellActive; /* [z][y][x] over the box, only when !allActive */
    size_t                      nActive = 0;
    uint64_t                    nextUid = 1;
    ~bGridState()
    {
        for (size_t d = 0; d < parts.size(); ++d) {
            if (backend.runtime() == Runtime::stream) {
                cudaSetDevice(backend.devId(int(d)));
                cudaFree(parts[d].infoDev);
                cudaFree(parts[d].activeMaskDev);
            }
        }
    }
};

}  // namespace detail

/* a cell of a block-sparse partition: block + position inside the 8 x 8 x 8 tile (the reference's bIndex, bIndex.h) */
struct bIdx
{
    uint32_t blk = 0;
    int32_t  x = 0, y = 0, z = 0;
};

/* the blocks a generic container visits for one data view, and what decides whether a cell of them exists
 * (the reference's bSpan, bSpan_imp.h:7-21: block id from the CUDA block, cell from the thread, active bit from the mask) */
struct bSpan
{
    using Idx = bIdx;
    uint32_t        first[2] = {0, 0}, count[2] = {0, 0}; /* up to two block ranges: BOUNDARY = lowest + highest layer */
    const uint32_t* info = nullptr;                        /* [block][32]: 27 neighbour ids + origin */
    const uint32_t* activeMask = nullptr;                  /* [block][16] or null: every cell inside the box is active */
    int32_t         gnx = 0, gny = 0, gnz = 0;
    uint32_t        nBlocksView() const { return count[0] + count[1]; }
    NEON_CUDA_HOST_DEVICE bool setAndValidate(bIdx& idx, uint32_t viewBlock, int x, int y, int z) const
    {
        if (viewBlock >= count[0] + count[1]) {
            return false;
        }
        idx.blk = viewBlock < count[0] ? first[0] + viewBlock : first[1] + (viewBlock - count[0]);
        idx.x = x;
        idx.y = y;
        idx.z = z;
        const uint32_t* line = info + size_t(idx.blk) * 32;
        if (int(line[27]) + x >= gnx || int(line[28]) + y >= gny || int(line[29]) + z >= gnz) {
            return false;
        }
        const int c = (z * detail::kB + y) * detail::kB + x;
        return activeMask == nullptr || ((activeMask[size_t(idx.blk) * 16 + (c >> 5)] >> (c & 31)) & 1u);
    }
};

class bGrid
{
   public:
    template <typename T, int C = 0>
    using Field = bField<T, C>;
    using Span = bSpan;
    using Idx = bIdx;
    static constexpr int blockEdge = detail::kB;

    bGrid() = default;

    template <typename ActiveCellLambda>
    bGrid(const Backend& bk, const index_3d& dim, ActiveCellLambda activeCellLambda, const domain::Stencil& stencil)
        : mS(std::make_shared<detail::bGridState>())
    {
        using namespace detail;
        auto& s = *mS;
        s.backend = bk;
        s.dim = dim;
        s.stencil = stencil;
        s.nb = index_3d((dim.x + kB - 1) / kB, (dim.y + kB - 1) / kB, (dim.z + kB - 1) / kB);
        if (dim.x <= 0 || dim.y <= 0 || dim.z <= 0) {
            NEON_THROW_UNSUPPORTED_OPERATION("empty box");
        }
        if (stencil.getRadius() > 1) {
            NEON_THROW_UNSUPPORTED_OPERATION("the LBM path serves radius-1 stencils (D3Q19 / D3Q27)");
        }
        const int nParts = bk.getDeviceCount();
        if (nParts > 1 && s.nb.z / nParts < 2) {
            NeonException e("bGrid");
            e << "every partition needs at least two block layers (" << s.nb.z << " layers over " << nParts << " devices)";
            NEON_THROW(e);
        }
        /* cell and block activity */
        const size_t cells = dim.rMul<size_t>();
        s.cellActive.assign(cells, 1);
        std::vector<uint8_t> blockActive(s.nb.rMul<size_t>(), 0);
        size_t               nActive = 0;
        bool                 all = true;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : nActive) reduction(&& : all)
#endif
        for (int z = 0; z < dim.z; ++z) {
            for (int y = 0; y < dim.y; ++y) {
                for (int x = 0; x < dim.x; ++x) {
                    const bool a = activeCellLambda(index_3d(x, y, z));
                    s.cellActive[(size_t(z) * dim.y + y) * dim.x + x] = a;
                    nActive += a;
                    all = all && a;
                    if (a) {
                        blockActive[(size_t(z / kB) * s.nb.y + y / kB) * s.nb.x + x / kB] = 1; /* benign race: only ever set to 1 */
                    }
                }
            }
        }
        s.nActive = nActive;
        s.allActive = all;
        if (all) {
            s.cellActive.clear();
            s.cellActive.shrink_to_fit();
        }
        /* partitions: block layers split evenly; local blocks sorted (bz, by, bx); ghosts = facing layers of the neighbours */
        auto layerBlocks = [&](int lz, std::vector<std::array<int, 3>>& out) {
            uint32_t n = 0;
            for (int by = 0; by < s.nb.y; ++by) {
                for (int bx = 0; bx < s.nb.x; ++bx) {
                    if (blockActive[(size_t(lz) * s.nb.y + by) * s.nb.x + bx]) {
                        out.push_back({lz, by, bx});
                        ++n;
                    }
                }
            }
            return n;
        };
        const int base = s.nb.z / nParts, rem = s.nb.z % nParts;
        int       l0 = 0;
        s.parts.resize(nParts);
        for (int d = 0; d < nParts; ++d) {
            const int l1 = l0 + base + (d < rem ? 1 : 0);
            auto&     p = s.parts[d];
            for (int lz = l0; lz < l1; ++lz) {
                const uint32_t n = layerBlocks(lz, p.coords);
                if (nParts > 1 && lz == l0 && d > 0) {
                    p.nDown = n;
                }
                if (nParts > 1 && lz == l1 - 1 && d < nParts - 1) {
                    p.nUp = n;
                }
            }
            p.nBlocks = uint32_t(p.coords.size());
            if (nParts > 1 && d > 0) {
                p.nGhostDown = layerBlocks(l0 - 1, p.coords);
            }
            if (nParts > 1 && d < nParts - 1) {
                p.nGhostUp = layerBlocks(l1, p.coords);
            }
            p.nAlloc = uint32_t(p.coords.size());
            l0 = l1;
            /* info lines: 27 neighbour ids + origin */
            std::vector<int64_t> lut(size_t(s.nb.z + 2) * (s.nb.y + 2) * (s.nb.x + 2), -1);
            auto                 lutAt = [&](int bz, int by, int bx) -> int64_t& {
                return lut[(size_t(bz + 1) * (s.nb.y + 2) + (by + 1)) * (s.nb.x + 2) + (bx + 1)];
            };
            for (uint32_t b = 0; b < p.nAlloc; ++b) {
                lutAt(p.coords[b][0], p.coords[b][1], p.coords[b][2]) = b;
            }
            std::vector<uint32_t> info(size_t(std::max<uint32_t>(p.nAlloc, 1)) * 32, 0);
            for (uint32_t b = 0; b < p.nAlloc; ++b) {
                uint32_t* line = info.data() + size_t(b) * 32;
                for (int k = 0; k < 27; ++k) {
                    line[k] = NLBM_NO_BLOCK;
                }
                if (b < p.nBlocks) {
                    for (int dz = -1; dz <= 1; ++dz) {
                        for (int dy = -1; dy <= 1; ++dy) {
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int64_t id = lutAt(p.coords[b][0] + dz, p.coords[b][1] + dy, p.coords[b][2] + dx);
                                line[(dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)] = id < 0 ? NLBM_NO_BLOCK : uint32_t(id);
                            }
                        }
                    }
                }
                line[27] = uint32_t(p.coords[b][2] * kB);
                line[28] = uint32_t(p.coords[b][1] * kB);
                line[29] = uint32_t(p.coords[b][0] * kB);
            }
            if (bk.runtime() == Runtime::stream) {
                bk.setDevice(d);
                NEON_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&p.infoDev), info.size() * 4));
                NEON_CUDA_CHECK(cudaMemcpy(p.infoDev, info.data(), info.size() * 4, cudaMemcpyHostToDevice));
                if (!all) {
                    std::vector<uint32_t> mask(size_t(std::max<uint32_t>(p.nAlloc, 1)) * 16, 0);
                    for (uint32_t b = 0; b < p.nAlloc; ++b) {
                        for (int c = 0; c < kBlockCells; ++c) {
                            const index_3d g(p.coords[b][2] * kB + (c & 7), p.coords[b][1] * kB + ((c >> 3) & 7), p.coords[b][0] * kB + (c >> 6));
                            if (isInsideDomain(g) && s.cellActive[(size_t(g.z) * dim.y + g.y) * dim.x + g.x]) {
                                mask[size_t(b) * 16 + (c >> 5)] |= 1u << (c & 31);
                            }
                        }
                    }
                    NEON_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&p.activeMaskDev), mask.size() * 4));
                    NEON_CUDA_CHECK(cudaMemcpy(p.activeMaskDev, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice));
                }
            }
        }
    }

    const Backend&         getBackend() const { return mS->backend; }
    const index_3d&        getDimension() const { return mS->dim; }
    const domain::Stencil& getStencil() const { return mS->stencil; }
    int                    getNumPartitions() const { return int(mS->parts.size()); }
    size_t                 getNumActiveCells() const { return mS->nActive; }
    size_t                 getNumBlocks() const
    {
        size_t n = 0;
        for (const auto& p : mS->parts) {
            n += p.nBlocks;
        }
        return n;
    }
    bool isInsideDomain(const index_3d& p) const
    {
        return p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < mS->dim.x && p.y < mS->dim.y && p.z < mS->dim.z;
    }
    bool isActive(const index_3d& p) const
    {
        return isInsideDomain(p) && (mS->allActive || mS->cellActive[(size_t(p.z) * mS->dim.y + p.y) * mS->dim.x + p.x]);
    }
    const detail::bPartitionInfo& partition(int setIdx) const { return mS->parts.at(setIdx); }
    const uint32_t*               activeMaskDev(int setIdx) const { return mS->parts.at(setIdx).activeMaskDev; }

    /* 19 or 27 if the grid's stencil is one of the two lattices of the kernel library in ITS order, else 0 (same rule as dGrid) */
    int latticeQ() const { return detail::latticeOf(mS->stencil.points()); }

    /* partition descriptor without field pointers */
    nlbm_block_desc descOf(int setIdx) const
    {
        const auto&     p = mS->parts.at(setIdx);
        nlbm_block_desc d{};
        d.info = p.infoDev;
        d.n_blocks = p.nBlocks;
        d.n_blocks_alloc = p.nAlloc;
        d.n_down = p.nDown;
        d.n_up = p.nUp;
        d.gnx = mS->dim.x;
        d.gny = mS->dim.y;
        d.gnz = mS->dim.z;
        return d;
    }

    template <typename T, int C = 0>
    bField<T, C> newField(const std::string& name, int cardinality, T outsideValue = T()) const
    {
        return bField<T, C>(*this, name, cardinality, outsideValue, mS->nextUid++);
    }

    /* blocks of a data view: STANDARD every local block, BOUNDARY the lowest and highest block layer of the partition (the
     * ones whose faces the halo update moves), INTERNAL the rest — the block ranges the LBM step kernel uses */
    bSpan getSpan(int setIdx, DataView dataView) const
    {
        const auto& p = mS->parts.at(setIdx);
        bSpan       sp;
        sp.info = p.infoDev;
        sp.activeMask = p.activeMaskDev;
        sp.gnx = mS->dim.x;
        sp.gny = mS->dim.y;
        sp.gnz = mS->dim.z;
        if (dataView == DataView::STANDARD) {
            sp.count[0] = p.nBlocks;
        } else if (dataView == DataView::INTERNAL) {
            sp.first[0] = p.nDown;
            sp.count[0] = p.nBlocks - p.nDown - p.nUp;
        } else {
            sp.count[0] = p.nDown;
            sp.first[1] = p.nBlocks - p.nUp;
            sp.count[1] = p.nUp;
        }
        return sp;
    }

    /* Grid::newContainer(name, loadingLambda): per-cell device lambdas over the active cells of the blocks of a data view
     * (LambdaExecutor.h:105-116, bSpan_imp.h:7-21).  Needs nvcc: include "Neon/domain/GenericContainer.h" from a .cu file. */
    template <typename LoadingLambda>
    set::Container newContainer(const std::string& name, LoadingLambda loadingLambda) const;

   private:
    std::shared_ptr<detail::bGridState> mS;
};

namespace detail {

/* Halo update of one block-sparse field: the facing z-slice of every boundary block, ordered by events only
 * (same protocol as DenseHaloImpl). */
struct BlockHaloImpl : set::Container::Impl
{
    bGrid                        grid;
    std::vector<void*>           mem;
    int                          elemBytes = 4, cardinality = 1, latticeQ = 0;
    set::TransferMode            mode = set::TransferMode::get;
    std::vector<cudaEvent_t>     ready, done;
    std::vector<nlbm_block_desc> desc;

    void push(int src, int dst, int dir, cudaStream_t st)
    {
        const auto&    p = grid.partition(dst);
        const uint32_t firstGhost = dir > 0 ? p.nBlocks : p.nBlocks + p.nGhostDown;
        check(nlbm_block_halo_push(&desc[src], mem[src], &desc[dst], mem[dst], firstGhost, elemBytes, cardinality, latticeQ, dir, st),
              "nlbm_block_halo_push");
    }
    void run(int streamIdx, DataView) override
    {
        const int n = grid.getNumPartitions();
        if (n == 1) {
            return;
        }
        if (backend.runtime() != Runtime::stream) {
            NEON_THROW_UNSUPPORTED_OPERATION("halo updates need Runtime::stream");
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
                if (nbr >= 0 && nbr < n) {
                    NEON_CUDA_CHECK(cudaStreamWaitEvent(st, ready[nbr], 0));
                }
            }
            for (int nbr : {d - 1, d + 1}) {
                if (nbr < 0 || nbr >= n) {
                    continue;
                }
                if (mode == set::TransferMode::get) {
                    push(nbr, d, nbr < d ? +1 : -1, st);
                } else {
                    push(d, nbr, nbr > d ? +1 : -1, st);
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
    }
    void run(int, int, DataView) override
    {
        NEON_THROW_UNSUPPORTED_OPERATION("a halo update involves every device; run it without a SetIdx");
    }
};

}  // namespace detail

template <typename T, int C = 0>
class bField
{
    using Codec = domain::FlagWordCodec<T>;
    static constexpr bool kFlagWords = Codec::enabled;
    static_assert(kFlagWords || std::is_same_v<T, float> || std::is_same_v<T, double> || std::is_same_v<T, int32_t> ||
                      std::is_same_v<T, uint32_t>,
                  "bField: float, double, 32-bit integers, or a type with a FlagWordCodec");

   public:
    using Type = T;
    using Grid = bGrid;
    using DeviceType = std::conditional_t<kFlagWords, uint32_t, T>;

    /* the reference's bPartition (bPartition.h:154-160) without the per-cell accessors the C ABI replaces */
    struct Partition
    {
        using Idx = bIdx;
        using Type = DeviceType;
        DeviceType*     memory = nullptr;
        nlbm_block_desc desc{};
        int             card = 0;
        const uint32_t* activeMask = nullptr; /* [block][16] or null */
        DeviceType      outsideValue{};
        NEON_CUDA_HOST_DEVICE DeviceType* mem() const { return memory; }
        NEON_CUDA_HOST_DEVICE int         cardinality() const { return card; }

        /* per-cell accessors for generic device lambdas (bPartition_imp.h:97-124, 194-198, 340-358), on the layout of this
         * library: pop[c][blk][z][y][x], 64-bit offsets */
        NEON_CUDA_HOST_DEVICE size_t offset(uint32_t blk, int x, int y, int z, int c) const
        {
            return (size_t(c) * desc.n_blocks_alloc + blk) * detail::kBlockCells + size_t((z * detail::kB + y) * detail::kB + x);
        }
        NEON_CUDA_HOST_DEVICE DeviceType& operator()(const bIdx& i, int c) const { return memory[offset(i.blk, i.x, i.y, i.z, c)]; }
        NEON_CUDA_HOST_DEVICE index_3d    getGlobalIndex(const bIdx& i) const
        {
            const uint32_t* line = static_cast<const uint32_t*>(desc.info) + size_t(i.blk) * 32;
            return {int(line[27]) + i.x, int(line[28]) + i.y, int(line[29]) + i.z};
        }
        /* the neighbour cell: inside the same tile, or in the tile the block's connectivity line names (a ghost block at a
         * partition face); it exists if that tile exists, the cell lies inside the box and is active */
        NEON_CUDA_HOST_DEVICE bool locate(const bIdx& i, int dx, int dy, int dz, uint32_t& blk, int& x, int& y, int& z) const
        {
            x = i.x + dx;
            y = i.y + dy;
            z = i.z + dz;
            const int fx = x < 0 ? -1 : (x >= detail::kB ? 1 : 0), fy = y < 0 ? -1 : (y >= detail::kB ? 1 : 0),
                      fz = z < 0 ? -1 : (z >= detail::kB ? 1 : 0);
            blk = i.blk;
            const uint32_t* info = static_cast<const uint32_t*>(desc.info);
            if (fx | fy | fz) {
                blk = info[size_t(i.blk) * 32 + (fx + 1) + 3 * (fy + 1) + 9 * (fz + 1)];
                if (blk == NLBM_NO_BLOCK) {
                    return false;
                }
                x -= detail::kB * fx;
                y -= detail::kB * fy;
                z -= detail::kB * fz;
            }
            const uint32_t* line = info + size_t(blk) * 32;
            if (int(line[27]) + x >= desc.gnx || int(line[28]) + y >= desc.gny || int(line[29]) + z >= desc.gnz) {
                return false;
            }
            const int c = (z * detail::kB + y) * detail::kB + x;
            return activeMask == nullptr || ((activeMask[size_t(blk) * 16 + (c >> 5)] >> (c & 31)) & 1u);
        }
        NEON_CUDA_HOST_DEVICE domain::NghData<DeviceType> getNghData(const bIdx& i, const index_3d& off, int c) const
        {
            domain::NghData<DeviceType> r;
            uint32_t                    blk;
            int                         x, y, z;
            r.mIsValid = locate(i, off.x, off.y, off.z, blk, x, y, z);
            r.mData = r.mIsValid ? memory[offset(blk, x, y, z, c)] : outsideValue;
            return r;
        }
        template <int dx, int dy, int dz>
        NEON_CUDA_HOST_DEVICE domain::NghData<DeviceType> getNghData(const bIdx& i, int c) const
        {
            return getNghData(i, index_3d(dx, dy, dz), c);
        }
        template <int dx, int dy, int dz>
        NEON_CUDA_HOST_DEVICE DeviceType getNghData(const bIdx& i, int c, DeviceType alternative) const
        {
            uint32_t blk;
            int      x, y, z;
            return locate(i, dx, dy, dz, blk, x, y, z) ? memory[offset(blk, x, y, z, c)] : alternative;
        }
    };
    using Idx = bIdx;

    bField() = default;

    const std::string& getName() const { return mS->name; }
    uint64_t           getUid() const { return mS->uid; }
    int                getCardinality() const { return mS->cardinality; }
    const bGrid&       getGrid() const { return mS->grid; }
    const index_3d&    getDimension() const { return mS->grid.getDimension(); }
    const Backend&     getBackend() const { return mS->grid.getBackend(); }
    bool               isValid() const { return bool(mS); }
    Partition&         getPartition(int setIdx) { return mS->parts.at(setIdx); }
    const Partition&   getPartition(int setIdx) const { return mS->parts.at(setIdx); }

    T& getReference(const index_3d& p, int card)
    {
        ensureHost();
        return mS->host[hostOffset(p, card)];
    }
    T operator()(const index_3d& p, int card) const
    {
        ensureHost();
        return mS->grid.isActive(p) ? mS->host[hostOffset(p, card)] : mS->outside;
    }
    T* hostData()
    {
        ensureHost();
        return mS->host.data();
    }
    const T* hostData() const
    {
        ensureHost();
        return mS->host.data();
    }

    template <typename Fn>
    void forEachActiveCell(Fn fn, computeMode_t mode = computeMode_t::par)
    {
        ensureHost();
        const bGrid&   g = mS->grid;
        const index_3d dim = g.getDimension();
        const int      card = mS->cardinality;
        T*             host = mS->host.data();
        const size_t   cells = dim.rMul<size_t>();
        auto           plane = [&](int z) {
            for (int y = 0; y < dim.y; ++y) {
                for (int x = 0; x < dim.x; ++x) {
                    const index_3d p(x, y, z);
                    if (!g.isActive(p)) {
                        continue;
                    }
                    const size_t o = (size_t(z) * dim.y + y) * dim.x + x;
                    for (int c = 0; c < card; ++c) {
                        fn(p, c, host[size_t(c) * cells + o]);
                    }
                }
            }
        };
        if (mode == computeMode_t::par) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
            for (int z = 0; z < dim.z; ++z) {
                plane(z);
            }
        } else {
            for (int z = 0; z < dim.z; ++z) {
                plane(z);
            }
        }
    }

    void updateDeviceData(int streamIdx = Backend::mainStreamIdx) { transfer(streamIdx, true); }
    /* block-sparse fields have no x-face cache (dField::commitWalls); same calls so that code is generic over the grid */
    void  commitWalls(int = Backend::mainStreamIdx) {}
    void  invalidateWalls() const {}
    void* wallCachePtr(int) const { return nullptr; }
    void updateHostData(int streamIdx = Backend::mainStreamIdx) { transfer(streamIdx, false); }

    set::Container newHaloUpdate(set::StencilSemantic semantic, set::TransferMode mode, Execution execution = Execution::device) const
    {
        if (execution != Execution::device) {
            NEON_THROW_UNSUPPORTED_OPERATION("host-side halo update: there is no CPU path");
        }
        auto         impl = std::make_shared<detail::BlockHaloImpl>();
        const bGrid& g = mS->grid;
        impl->grid = g;
        impl->backend = g.getBackend();
        impl->kind = set::Container::Kind::halo;
        impl->mode = mode;
        impl->elemBytes = int(sizeof(DeviceType));
        impl->cardinality = mS->cardinality;
        if (semantic == set::StencilSemantic::streaming) {
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

    void ioToVtk(const std::string&, const std::string&, bool = false, IoFileType = IoFileType::ASCII, bool = false) const
    {
        NEON_THROW_UNSUPPORTED_OPERATION("VTK export of block-sparse fields (use dGrid for --visual runs)");
    }

   private:
    friend class bGrid;
    struct State
    {
        bGrid               grid;
        std::string         name;
        int                 cardinality = 0;
        T                   outside{};
        uint64_t            uid = 0;
        std::vector<T>      host;
        std::vector<void*>  dev;
        std::vector<Partition> parts;
        std::vector<DeviceType*> staging; /* pinned [cardinality][nAlloc][512] per partition */
        ~State()
        {
            const Backend& bk = grid.getBackend();
            for (size_t d = 0; d < dev.size(); ++d) {
                if (dev[d]) {
                    cudaSetDevice(bk.devId(int(d)));
                    cudaFree(dev[d]);
                }
                if (staging[d]) {
                    cudaFreeHost(staging[d]);
                }
            }
        }
    };

    bField(const bGrid& grid, const std::string& name, int cardinality, T outside, uint64_t uid) : mS(std::make_shared<State>())
    {
        if ((C != 0 && cardinality != C) || cardinality < 1 || cardinality > 27 || (kFlagWords && cardinality != 1)) {
            NeonException e("bField");
            e << "unsupported cardinality " << cardinality;
            NEON_THROW(e);
        }
        auto& s = *mS;
        s.grid = grid;
        s.name = name;
        s.cardinality = cardinality;
        s.outside = outside;
        s.uid = uid;
        /* host mirror and staging buffers are allocated at their first use (ensureHost) */
        const Backend& bk = grid.getBackend();
        for (int d = 0; d < grid.getNumPartitions(); ++d) {
            const size_t n = size_t(cardinality) * std::max<uint32_t>(grid.partition(d).nAlloc, 1) * detail::kBlockCells;
            void*        p = nullptr;
            DeviceType*  st = nullptr;
            if (bk.runtime() == Runtime::stream) {
                bk.setDevice(d);
                NEON_CUDA_CHECK(cudaMalloc(&p, n * sizeof(DeviceType)));
                NEON_CUDA_CHECK(cudaMemset(p, 0, n * sizeof(DeviceType)));
            }
            s.dev.push_back(p);
            s.staging.push_back(st);
            Partition part;
            part.memory = static_cast<DeviceType*>(p);
            part.desc = grid.descOf(d);
            part.card = cardinality;
            part.activeMask = grid.activeMaskDev(d);
            if constexpr (!kFlagWords) {
                part.outsideValue = outside;
            }
            s.parts.push_back(part);
        }
    }

    size_t hostOffset(const index_3d& p, int card) const
    {
        const index_3d& dim = mS->grid.getDimension();
        return (size_t(card) * dim.z + p.z) * size_t(dim.y) * dim.x + size_t(p.y) * dim.x + p.x;
    }

    void ensureHost() const
    {
        auto& s = *mS;
        if (!s.host.empty()) {
            return;
        }
        s.host.assign(s.grid.getDimension().template rMul<size_t>() * size_t(s.cardinality), s.outside);
        const Backend& bk = s.grid.getBackend();
        for (int d = 0; bk.runtime() == Runtime::stream && d < s.grid.getNumPartitions(); ++d) {
            const size_t n = size_t(s.cardinality) * std::max<uint32_t>(s.grid.partition(d).nAlloc, 1) * detail::kBlockCells;
            NEON_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&s.staging[d]), n * sizeof(DeviceType), cudaHostAllocDefault));
        }
    }

    static DeviceType toDevice(const T& v)
    {
        if constexpr (kFlagWords) {
            return Codec::pack(v);
        } else {
            return v;
        }
    }
    static T fromDevice(const DeviceType& v)
    {
        if constexpr (kFlagWords) {
            return Codec::unpack(v);
        } else {
            return v;
        }
    }

    void transfer(int streamIdx, bool toDev)
    {
        using namespace detail;
        const bGrid&   g = mS->grid;
        const Backend& bk = g.getBackend();
        ensureHost();
        if (bk.runtime() != Runtime::stream) {
            return;
        }
        const index_3d dim = g.getDimension();
        const size_t   cells = dim.rMul<size_t>();
        const int      card = mS->cardinality;
        /* cells of a block that are outside the box or inactive */
        const DeviceType hole = kFlagWords ? DeviceType(uint32_t(NLBM_UNDEFINED) << NLBM_FLAG_CLASS_SHIFT) : DeviceType(0);
        for (int d = 0; d < g.getNumPartitions(); ++d) {
            const auto&    part = g.partition(d);
            const uint32_t nBlk = toDev ? part.nAlloc : part.nBlocks;
            const size_t   perComp = size_t(std::max<uint32_t>(part.nAlloc, 1)) * kBlockCells;
            DeviceType*    st = mS->staging[d];
            bk.setDevice(d);
            cudaStream_t stream = bk.stream(d, streamIdx);
            const size_t bytes = size_t(card) * perComp * sizeof(DeviceType);
            if (!toDev) {
                NEON_CUDA_CHECK(cudaMemcpyAsync(st, mS->dev[d], bytes, cudaMemcpyDeviceToHost, stream));
                NEON_CUDA_CHECK(cudaStreamSynchronize(stream));
            }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
            for (int64_t b = 0; b < int64_t(nBlk); ++b) {
                const int bz = part.coords[b][0] * kB, by = part.coords[b][1] * kB, bx = part.coords[b][2] * kB;
                for (int c = 0; c < card; ++c) {
                    DeviceType* blk = st + size_t(c) * perComp + size_t(b) * kBlockCells;
                    T*          host = mS->host.data() + size_t(c) * cells;
                    for (int z = 0; z < kB; ++z) {
                        for (int y = 0; y < kB; ++y) {
                            for (int x = 0; x < kB; ++x) {
                                const index_3d p(bx + x, by + y, bz + z);
                                const bool     live = g.isActive(p);
                                DeviceType&    w = blk[z * 64 + y * 8 + x];
                                if (toDev) {
                                    w = live ? toDevice(host[(size_t(p.z) * dim.y + p.y) * dim.x + p.x]) : hole;
                                } else if (live) {
                                    host[(size_t(p.z) * dim.y + p.y) * dim.x + p.x] = fromDevice(w);
                                }
                            }
                        }
                    }
                }
            }
            if (toDev) {
                NEON_CUDA_CHECK(cudaMemcpyAsync(mS->dev[d], st, bytes, cudaMemcpyHostToDevice, stream));
            }
        }
    }

    std::shared_ptr<State> mS;
};

}  // namespace Neon
