// Lbm.h — the LBM hot path behind the names the reference benchmark uses.
//
//   CellType                        benchmarks/lbm-lid-driven-cavity-flow/src/CellType.h:1-40
//   D3Q19Template                   src/D3Q19.h:7-175 (c_vect :23-44, t_vect :112-132)
//   D3Q27Template                   apps/lbmMultiRes/lattice.h:15-77 (rest population first)
//   LbmContainers::iteration        src/LbmTools.h:285-325  -> nlbm_d3q{19,27}_*_dense_step
//   LbmContainers::computeWallNghMask  LbmTools.h:344-376   -> nlbm_dense_wall_mask
//   LbmContainers::computeRhoAndU   LbmTools.h:384-437      -> nlbm_d3q19_*_dense_rho_u
//   LbmIterationD3Q19               src/LbmIteration.h:19-101 (two pre-built Skeletons, run() flips the parity)
//
// In the reference these are application headers whose device lambdas run through Neon's generic launcher; here every
// container is a device-managed container whose body is one call into libneon_lbm.so per device.  Nothing here computes.
#pragma once

#include <ostream>
#include <type_traits>
#include <vector>

#include "Neon/Neon.h"
#include "Neon/domain/bGrid.h"
#include "Neon/domain/dGrid.h"
#include "Neon/set/Container.h"
#include "Neon/skeleton/Skeleton.h"

struct CellType
{
    enum Classification : int
    {
        bounceBack = NLBM_BOUNCE_BACK,
        movingWall = NLBM_MOVING_WALL,
        bulk = NLBM_BULK,
        undefined = NLBM_UNDEFINED
    };
    /* the reference's default constructor ignores its argument and yields a bulk cell (CellType.h:13-18) */
    CellType(int = 0) : wallNghBitflag(0), classification(bulk) {}
    explicit CellType(Classification c, uint32_t n = 0) : wallNghBitflag(n), classification(c) {}
    uint32_t       wallNghBitflag;
    Classification classification;
};
inline std::ostream& operator<<(std::ostream& os, const CellType& c)
{
    return os << static_cast<double>(c.classification);
}

namespace Neon::domain {
/* device representation of CellType: the 32-bit flag word of include/neon_lbm.h */
template <>
struct FlagWordCodec<CellType, void>
{
    static constexpr bool enabled = true;
    static uint32_t       pack(const CellType& c)
    {
        return (c.wallNghBitflag & NLBM_FLAG_MASK_BITS) | (uint32_t(c.classification) << NLBM_FLAG_CLASS_SHIFT);
    }
    static CellType unpack(uint32_t w)
    {
        return CellType(static_cast<CellType::Classification>(NLBM_FLAG_CLASS(w)), w & NLBM_FLAG_MASK_BITS);
    }
};
}  // namespace Neon::domain

template <typename StorageFP, typename ComputeFP>
struct D3Q19Template
{
    static constexpr int Q = 19;
    static constexpr int D = 3;
    static constexpr int centerDirection = 9;
    static constexpr int goRangeBegin = 0, goRangeEnd = 8, goBackOffset = 10;

    explicit D3Q19Template(const Neon::Backend&)
    {
        /* direction k and k + 10 are opposite; 9 is the rest population */
        const int go[9][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {0, -1, -1}, {0, -1, 1}};
        c_vect.resize(Q);
        t_vect.resize(Q);
        opp_vect.resize(Q);
        for (int k = 0; k < 9; ++k) {
            c_vect[k] = Neon::index_3d(go[k][0], go[k][1], go[k][2]);
            c_vect[k + goBackOffset] = -c_vect[k];
            const double w = k < 3 ? 1. / 18. : 1. / 36.;
            t_vect[k] = t_vect[k + goBackOffset] = w;
            opp_vect[k] = k + goBackOffset;
            opp_vect[k + goBackOffset] = k;
        }
        c_vect[centerDirection] = Neon::index_3d(0, 0, 0);
        t_vect[centerDirection] = 1. / 3.;
        opp_vect[centerDirection] = centerDirection;
    }
    template <int go>
    static constexpr int getOpposite()
    {
        return go == centerDirection ? centerDirection : go <= goRangeEnd ? go + goBackOffset : go - goBackOffset;
    }
    std::vector<double>         t_vect;
    std::vector<Neon::index_3d> c_vect;
    std::vector<int>            opp_vect;
};

template <typename StorageFP, typename ComputeFP>
struct D3Q27Template
{
    static constexpr int Q = 27;
    static constexpr int D = 3;
    static constexpr int centerDirection = 0;

    explicit D3Q27Template(const Neon::Backend&)
    {
        /* x slowest (0, -1, +1), then y, then z, each in the order 0, -1, +1 */
        const int order[3] = {0, -1, 1};
        for (int ix : order) {
            for (int iy : order) {
                for (int iz : order) {
                    c_vect.emplace_back(ix, iy, iz);
                    const int nz = (ix != 0) + (iy != 0) + (iz != 0);
                    t_vect.push_back(nz == 0 ? 8. / 27. : nz == 1 ? 2. / 27. : nz == 2 ? 1. / 54. : 1. / 216.);
                }
            }
        }
        opp_vect.resize(Q);
        for (int k = 0; k < Q; ++k) {
            for (int j = 0; j < Q; ++j) {
                if (c_vect[j] == -c_vect[k]) {
                    opp_vect[k] = j;
                }
            }
        }
    }
    std::vector<double>         t_vect;
    std::vector<Neon::index_3d> c_vect;
    std::vector<int>            opp_vect;
};

namespace Neon::lbm {

/* process-wide options of the kernel library for containers built afterwards: arithmetic mode (NLBM_ARITH_FAST:
 * FMAs in the storage precision, within the north-star tolerance; NLBM_ARITH_REFERENCE: the reference's rounding bit
 * for bit) and tuning bits (NLBM_OPT_*).  They never change which cells are written. */
inline int& kernelOptions()
{
    static int opts = NLBM_ARITH_FAST;
    return opts;
}

}  // namespace Neon::lbm

template <typename Lattice, typename PopulationField, typename LbmComputeType>
struct LbmContainers
{
    using LbmStoreType = typename PopulationField::Type;
    using CellTypeField = typename PopulationField::Grid::template Field<CellType, 1>;
    using Rho = typename PopulationField::Grid::template Field<LbmStoreType, 1>;
    using U = typename PopulationField::Grid::template Field<LbmStoreType, 3>;

    /* One fused pull-stream + BGK collide over the cells of the data view.  fIn: const STENCIL read with the given
     * halo semantic, fOut: MAP write, flags: MAP read (LbmTools.h:296-299). */
    static auto iteration(Neon::set::StencilSemantic stencilSemantic, const PopulationField& fInField, const CellTypeField& cellTypeField,
                          const LbmComputeType omega, PopulationField& fOutField) -> Neon::set::Container
    {
        constexpr bool isBlock = std::is_same_v<typename PopulationField::Grid, Neon::bGrid>;
        using Desc = std::conditional_t<isBlock, nlbm_block_desc, nlbm_dense_desc>;
        using StepFn = int (*)(const Desc*, double, int, int, void*);
        constexpr bool f32 = std::is_same_v<LbmStoreType, float> && std::is_same_v<LbmComputeType, float>;
        constexpr bool f64 = std::is_same_v<LbmStoreType, double> && std::is_same_v<LbmComputeType, double>;
        constexpr bool f32c64 = std::is_same_v<LbmStoreType, float> && std::is_same_v<LbmComputeType, double>;
        StepFn         step = nullptr;
        if constexpr (isBlock) {
            if constexpr (Lattice::Q == 19 && f32) {
                step = nlbm_d3q19_f32_block_step;
            } else if constexpr (Lattice::Q == 19 && f64) {
                step = nlbm_d3q19_f64_block_step;
            } else if constexpr (Lattice::Q == 27 && f32) {
                step = nlbm_d3q27_f32_block_step;
            } else if constexpr (Lattice::Q == 27 && f64) {
                step = nlbm_d3q27_f64_block_step;
            }
        } else {
            if constexpr (Lattice::Q == 19 && f32) {
                step = nlbm_d3q19_f32_dense_step;
            } else if constexpr (Lattice::Q == 19 && f32c64) {
                step = nlbm_d3q19_f32c64_dense_step;
            } else if constexpr (Lattice::Q == 19 && f64) {
                step = nlbm_d3q19_f64_dense_step;
            } else if constexpr (Lattice::Q == 27 && f32) {
                step = nlbm_d3q27_f32_dense_step;
            } else if constexpr (Lattice::Q == 27 && f64) {
                step = nlbm_d3q27_f64_dense_step;
            }
        }
        if (step == nullptr) {
            NEON_THROW_UNSUPPORTED_OPERATION("store/compute precision pair (dGrid: f/f, f/d (D3Q19), d/d; bGrid: f/f, d/d)");
        }
        if (fInField.getUid() == fOutField.getUid()) {
            NEON_THROW_UNSUPPORTED_OPERATION("the pull scheme needs two population fields (LbmIteration.h:38-39)");
        }
        if (fInField.getCardinality() != Lattice::Q || fOutField.getCardinality() != Lattice::Q) {
            NEON_THROW_UNSUPPORTED_OPERATION("population fields must have the lattice's cardinality");
        }
        const Neon::Backend bk = fInField.getBackend();
        const int           opts = Neon::lbm::kernelOptions();
        const double        om = static_cast<double>(omega);
        return Neon::set::Container::factoryDeviceManaged(
            "LBM_iteration_D3Q" + std::to_string(Lattice::Q), bk, [&](Neon::SetIdx setIdx, Neon::set::Loader& L) {
                auto            fIn = L.load(fInField, Neon::Pattern::STENCIL, stencilSemantic);
                auto            fOut = L.load(fOutField);
                auto            flg = L.load(cellTypeField);
                Desc            d = fIn.desc;
                d.pop_in = fIn.mem();
                d.pop_out = fOut.mem();
                d.flags = flg.mem();
                const int             dev = setIdx.idx;
                const PopulationField outHandle = fOutField; /* shallow handle: its x-face cache may be (in)validated later */
                return [=](int streamIdx, Neon::DataView dataView) {
                    Desc dd = d;
                    if constexpr (!isBlock) {
                        dd.wall_cache = outHandle.wallCachePtr(dev);
                    }
                    Neon::detail::check(step(&dd, om, static_cast<int>(dataView), opts, bk.stream(dev, streamIdx)), "nlbm step");
                };
            });
    }

    /* LbmTools.h:344-376, run in place as RunCavityTwoPop.cu:239 does.  The launcher checks the library's count of
     * bulk-cell neighbours that fall outside the domain (the reference would read invalid data there). */
    static auto computeWallNghMask(const CellTypeField& infoInField, CellTypeField& infoOutpeField) -> Neon::set::Container
    {
        if (infoInField.getUid() != infoOutpeField.getUid()) {
            NEON_THROW_UNSUPPORTED_OPERATION("the wall mask is built in place (RunCavityTwoPop.cu:239)");
        }
        const Neon::Backend bk = infoInField.getBackend();
        /* one error counter per device, allocated once with the container (device word + pinned host word): a run enqueues
         * memset, kernel and read-back on every device first and only then waits — no allocation, no per-device host sync */
        struct Counters
        {
            Neon::Backend         bk;
            std::vector<int32_t*> dev, host;
            ~Counters()
            {
                for (size_t d = 0; d < dev.size(); ++d) {
                    bk.setDevice(int(d));
                    if (dev[d]) {
                        cudaFree(dev[d]);
                    }
                    if (host[d]) {
                        cudaFreeHost(host[d]);
                    }
                }
            }
        };
        auto cnt = std::make_shared<Counters>();
        cnt->bk = bk;
        cnt->dev.assign(bk.getDeviceCount(), nullptr);
        cnt->host.assign(bk.getDeviceCount(), nullptr);
        for (int d = 0; d < bk.getDeviceCount(); ++d) {
            bk.setDevice(d);
            NEON_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&cnt->dev[d]), sizeof(int32_t)));
            NEON_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&cnt->host[d]), sizeof(int32_t)));
            *cnt->host[d] = 0;
        }
        auto check = [cnt, bk](int dev, int streamIdx) {
            bk.setDevice(dev);
            NEON_CUDA_CHECK(cudaStreamSynchronize(bk.stream(dev, streamIdx)));
            const int32_t nBad = *cnt->host[dev];
            if (nBad != 0) {
                Neon::NeonException e("computeWallNghMask");
                e << nBad << " bulk-cell neighbours fall outside the domain: enclose the geometry with non-bulk cells";
                NEON_THROW(e);
            }
        };
        auto pending = std::make_shared<std::vector<char>>(bk.getDeviceCount(), 0);
        auto c = Neon::set::Container::factoryDeviceManaged(
            "LBM_computeWallNghMask", bk, [&](Neon::SetIdx setIdx, Neon::set::Loader& L) {
                auto in = L.load(infoInField, Neon::Pattern::STENCIL);
                auto out = L.load(infoOutpeField);
                auto d = out.desc; /* nlbm_dense_desc or nlbm_block_desc */
                d.flags = out.mem();
                (void)in;
                const int dev = setIdx.idx;
                return [=](int streamIdx, Neon::DataView) {
                    int32_t*     bad = cnt->dev[dev];
                    cudaStream_t st = bk.stream(dev, streamIdx);
                    NEON_CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
                    if constexpr (std::is_same_v<decltype(d), nlbm_block_desc>) {
                        Neon::detail::check(nlbm_block_wall_mask(&d, Lattice::Q, bad, st), "nlbm_block_wall_mask");
                    } else {
                        Neon::detail::check(nlbm_dense_wall_mask(&d, Lattice::Q, bad, st), "nlbm_dense_wall_mask");
                    }
                    NEON_CUDA_CHECK(cudaMemcpyAsync(cnt->host[dev], bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                    (*pending)[dev] = 1;
                };
            });
        /* the geometry check needs the counters on the host: after ALL devices were issued (all-device run) */
        c.template as<Neon::set::detail::DeviceManagedImpl>()->afterAll = [=](int streamIdx) {
            for (int d = 0; d < int(pending->size()); ++d) {
                if ((*pending)[d]) {
                    (*pending)[d] = 0;
                    check(d, streamIdx);
                }
            }
        };
        return c;
    }

    /* LbmTools.h:384-437 (D3Q19): rho and u of every cell from the populations */
    static auto computeRhoAndU(const PopulationField& fInField, const CellTypeField& cellTypeField, Rho& rhoField, U& uField)
        -> Neon::set::Container
    {
        using Fn = int (*)(const nlbm_dense_desc*, void*, void*, void*);
        Fn fn = nullptr;
        if constexpr (std::is_same_v<typename PopulationField::Grid, Neon::bGrid>) {
            NEON_THROW_UNSUPPORTED_OPERATION("computeRhoAndU on bGrid (the library exports the dense variant; use dGrid for --visual)");
        } else if constexpr (Lattice::Q == 19 && std::is_same_v<LbmStoreType, float>) {
            fn = nlbm_d3q19_f32_dense_rho_u;
        } else if constexpr (Lattice::Q == 19 && std::is_same_v<LbmStoreType, double>) {
            fn = nlbm_d3q19_f64_dense_rho_u;
        } else {
            NEON_THROW_UNSUPPORTED_OPERATION("computeRhoAndU is a D3Q19 container (LbmTools.h:384-437)");
        }
        const Neon::Backend bk = fInField.getBackend();
        return Neon::set::Container::factoryDeviceManaged(
            "LBM_computeRhoAndU", bk, [&](Neon::SetIdx setIdx, Neon::set::Loader& L) {
                auto            fIn = L.load(fInField, Neon::Pattern::STENCIL);
                auto            flg = L.load(cellTypeField);
                auto            rho = L.load(rhoField);
                auto      u = L.load(uField);
                const int dev = setIdx.idx;
                if constexpr (std::is_same_v<typename PopulationField::Grid, Neon::bGrid>) {
                    return std::function<void(int, Neon::DataView)>();
                } else {
                    nlbm_dense_desc d = fIn.desc;
                    d.pop_in = fIn.mem();
                    d.flags = flg.mem();
                    return std::function<void(int, Neon::DataView)>([=](int streamIdx, Neon::DataView) {
                        Neon::detail::check(fn(&d, rho.mem(), u.mem(), bk.stream(dev, streamIdx)), "nlbm dense rho_u");
                    });
                }
            });
    }
};

/* LbmIteration.h:19-101, for both lattices: skeleton 0 streams pop0 -> pop1, skeleton 1 pop1 -> pop0 */
template <typename Lattice, typename PopulationField, typename LbmComputeType>
struct LbmIterationT
{
    using LbmStoreType = typename PopulationField::Type;
    using CellTypeField = typename PopulationField::Grid::template Field<CellType, 1>;
    using LbmTools = LbmContainers<Lattice, PopulationField, LbmComputeType>;

    LbmIterationT(Neon::set::StencilSemantic stencilSemantic, Neon::skeleton::Occ occ, Neon::set::TransferMode transfer,
                  PopulationField& fIn, PopulationField& fOut, CellTypeField& cellTypeField, LbmComputeType omega, bool cudaGraph = false)
    {
        pop[0] = fIn;
        pop[1] = fOut;
        flag = cellTypeField;
        mOmega = static_cast<double>(omega);
        for (int target = 0; target < 2; ++target) {
            std::vector<Neon::set::Container> ops;
            ops.push_back(LbmTools::iteration(stencilSemantic, pop[target], cellTypeField, omega, pop[1 - target]));
            lbmTwoPop[target] = Neon::skeleton::Skeleton(fIn.getBackend());
            lbmTwoPop[target].sequence(ops, "LBM_iteration_" + std::to_string(target), Neon::skeleton::Options(occ, transfer, cudaGraph));
        }
    }
    auto getInput() -> PopulationField& { return pop[parity]; }
    auto getOutput() -> PopulationField& { return pop[1 - parity]; }
    auto run() -> void
    {
        lbmTwoPop[parity].run();
        parity = 1 - parity;
    }
    /* `n` iterations with ONE library call (nlbm_dense_step_n) when the field is a single dense partition: a chain of
     * dependent launches whose tiles wait plane-wise for the previous iteration, so that launch gap, ramp-up and tail of
     * consecutive iterations overlap (small boxes).  Anything else runs n x run(). */
    auto runMany(int n) -> void
    {
        constexpr bool isBlock = std::is_same_v<typename PopulationField::Grid, Neon::bGrid>;
        if constexpr (!isBlock) {
            constexpr bool f32 = std::is_same_v<LbmStoreType, float> && std::is_same_v<LbmComputeType, float>;
            constexpr bool f64 = std::is_same_v<LbmStoreType, double> && std::is_same_v<LbmComputeType, double>;
            constexpr bool f32c64 = std::is_same_v<LbmStoreType, float> && std::is_same_v<LbmComputeType, double>;
            constexpr int  kind = Lattice::Q == 19 ? (f32 ? 0 : (f64 ? 1 : (f32c64 ? 2 : -1))) : (f32 ? 3 : (f64 ? 4 : -1));
            const Neon::Backend& bk = pop[0].getBackend();
            if (n > 0 && kind >= 0 && bk.getDeviceCount() == 1 && pop[0].getGrid().zHalo() == 0) {
                nlbm_dense_desc d = pop[parity].getPartition(0).desc;
                d.pop_in = pop[parity].getPartition(0).mem();
                d.pop_out = pop[1 - parity].getPartition(0).mem();
                d.flags = flag.getPartition(0).mem();
                d.wall_cache = pop[1 - parity].wallCachePtr(0);
                Neon::detail::check(nlbm_dense_step_n(kind, &d, pop[parity].wallCachePtr(0), mOmega, n, Neon::lbm::kernelOptions(),
                                                      bk.stream(0, Neon::Backend::mainStreamIdx)),
                                    "nlbm_dense_step_n");
                parity = (parity + n) & 1;
                return;
            }
        }
        for (int i = 0; i < n; ++i) {
            run();
        }
    }
    /* whether runMany beats n x run() for this field (measured on B200): one dense partition of more than one
     * chip-load of blocks (~400 000 cells) up to 2^24 cells */
    auto chainPays() const -> bool
    {
        if constexpr (std::is_same_v<typename PopulationField::Grid, Neon::bGrid>) {
            return false;
        } else {
            const auto   dim = pop[0].getDimension();
            const size_t cells = dim.template rMul<size_t>();
            return pop[0].getBackend().getDeviceCount() == 1 && pop[0].getGrid().zHalo() == 0 && cells > 400000 && cells <= (size_t(1) << 24);
        }
    }
    auto sync() -> void { pop[0].getBackend().syncAll(); }
    auto skeleton(int target) -> Neon::skeleton::Skeleton& { return lbmTwoPop[target]; }

   private:
    Neon::skeleton::Skeleton lbmTwoPop[2];
    PopulationField          pop[2];
    CellTypeField            flag;
    double                   mOmega = 0;
    int                      parity = 0;
};

template <typename PopulationField, typename LbmComputeType>
using LbmIterationD3Q19 = LbmIterationT<D3Q19Template<typename PopulationField::Type, LbmComputeType>, PopulationField, LbmComputeType>;
template <typename PopulationField, typename LbmComputeType>
using LbmIterationD3Q27 = LbmIterationT<D3Q27Template<typename PopulationField::Type, LbmComputeType>, PopulationField, LbmComputeType>;
