// Skeleton.h — schedules a sequence of Containers on the Backend's stream sets, inserts the halo updates the declared
// stencil reads need and overlaps them with computation (OCC).
//
// Mirrors libNeonSkeleton: Skeleton::sequence / run (include/Neon/skeleton/Skeleton.h:32-65), Options(Occ, TransferMode),
// Occ (Occ.h:8-14), and what MultiXpuGraph does for a sequence (src/skeleton/internal/multiGpuGraph.cpp): dependencies
// from the Loader tokens (:43-70), OCC split of a stencil node into INTERNAL + BOUNDARY clones (:120-301), a halo-update
// node on every stencil-read edge whose consumer is not INTERNAL (:304-352), stream mapping and event insertion
// (libNeonSet/src/set/container/Graph.cpp:690-838), sequential host issue (:992-1030).
//
// Schedule produced per stencil container with more than one device:
//   Occ::none      stream 0: halo update -> STANDARD
//   Occ::standard  stream 0: INTERNAL            || stream 1: halo update -> BOUNDARY        (fork/join by events)
// Differences by design: ordering is by CUDA events only (the reference's halo update blocks the host on every device,
// SynchronizationContainer.h:37-42); the BOUNDARY view is z_local in {0, nz-1} (the reference folds it onto {0,1},
// SURVEY.md fact 7); with one device a whole run() can be captured once into a CUDA graph and replayed (the launches are
// then frozen: call sequence() again after anything the containers look up at launch time changed, e.g. a field's x-face cache).
#pragma once

#include <fstream>
#include <string>
#include <vector>

#include "Neon/set/Backend.h"
#include "Neon/set/Container.h"

namespace Neon::skeleton {

enum class Occ
{
    standard,
    extended,
    twoWayExtended,
    none
};
struct OccUtils
{
    static std::string toString(Occ occ)
    {
        switch (occ) {
            case Occ::standard: return "standard";
            case Occ::extended: return "extended";
            case Occ::twoWayExtended: return "twoWayExtended";
            default: return "none";
        }
    }
    static Occ fromString(const std::string& s)
    {
        for (Occ o : {Occ::standard, Occ::extended, Occ::twoWayExtended, Occ::none}) {
            if (toString(o) == s) {
                return o;
            }
        }
        NeonException e("OccUtils::fromString");
        e << "unknown OCC option " << s;
        NEON_THROW(e);
    }
};

class Options
{
   public:
    Options() = default;
    Options(Occ occ, set::TransferMode transferMode, bool cudaGraph = false)
        : mOcc(occ), mTransferMode(transferMode), mCudaGraph(cudaGraph)
    {
    }
    Occ               occ() const { return mOcc; }
    set::TransferMode transferMode() const { return mTransferMode; }
    bool              cudaGraph() const { return mCudaGraph; }

   private:
    Occ               mOcc = Occ::none;
    set::TransferMode mTransferMode = set::TransferMode::get;
    bool              mCudaGraph = false; /* one device only: replay the sequence as a CUDA graph */
};

class Skeleton
{
   public:
    struct Node
    {
        enum Kind
        {
            fork,
            join,
            halo,
            compute
        } kind;
        std::string    name;
        int            stream = 0;
        DataView       view = DataView::STANDARD;
        set::Container container;
    };

    Skeleton() = default;
    explicit Skeleton(const Backend& bk) : mBk(bk), mHasBk(true) {}
    ~Skeleton()
    {
        if (mGraphExec) {
            cudaGraphExecDestroy(mGraphExec);
        }
    }
    Skeleton(const Skeleton&) = delete;
    Skeleton& operator=(const Skeleton&) = delete;
    Skeleton(Skeleton&& o) noexcept { *this = std::move(o); }
    Skeleton& operator=(Skeleton&& o) noexcept
    {
        if (this == &o) {
            return *this;
        }
        if (mGraphExec) { /* a skeleton that already captured a graph is being re-assigned */
            cudaGraphExecDestroy(mGraphExec);
            mGraphExec = nullptr;
        }
        mBk = o.mBk;
        mHasBk = o.mHasBk;
        mName = std::move(o.mName);
        mOptions = o.mOptions;
        mNodes = std::move(o.mNodes);
        mForkEv = std::move(o.mForkEv);
        mJoinEv = std::move(o.mJoinEv);
        mGraphExec = o.mGraphExec;
        o.mGraphExec = nullptr;
        return *this;
    }

    void sequence(const std::vector<set::Container>& operations, const std::string& name = "", Options options = Options())
    {
        if (!mHasBk) {
            NeonException e("Skeleton::sequence");
            e << "skeleton without a backend";
            NEON_THROW(e);
        }
        mName = name;
        mOptions = options;
        mNodes.clear();
        const bool multi = mBk.getDeviceCount() > 1;
        for (const auto& c : operations) {
            std::vector<set::Container> halos;
            if (multi && c.getKind() == set::Container::Kind::compute) {
                for (const auto& t : c.getTokens()) {
                    if (t.access == set::Access::read && t.pattern == Pattern::STENCIL && t.newHaloUpdate) {
                        halos.push_back(t.newHaloUpdate(t.semantic, options.transferMode(), Execution::device));
                    }
                }
            }
            if (!halos.empty() && options.occ() != Occ::none) {
                mNodes.push_back({Node::fork, "fork", 0, DataView::STANDARD, {}});
                mNodes.push_back({Node::compute, c.getName(), 0, DataView::INTERNAL, c});
                for (auto& h : halos) {
                    mNodes.push_back({Node::halo, h.getName(), 1, DataView::STANDARD, h});
                }
                mNodes.push_back({Node::compute, c.getName(), 1, DataView::BOUNDARY, c});
                mNodes.push_back({Node::join, "join", 0, DataView::STANDARD, {}});
            } else {
                for (auto& h : halos) {
                    mNodes.push_back({Node::halo, h.getName(), 0, DataView::STANDARD, h});
                }
                mNodes.push_back({c.getKind() == set::Container::Kind::halo ? Node::halo : Node::compute, c.getName(), 0,
                                  DataView::STANDARD, c});
            }
        }
        int width = 1;
        for (const auto& n : mNodes) {
            width = std::max(width, n.stream + 1);
        }
        mBk.setAvailableStreamSet(width);
        if (mBk.runtime() == Runtime::stream && width > 1 && mForkEv.empty()) {
            for (int d = 0; d < mBk.getDeviceCount(); ++d) {
                mForkEv.push_back(mBk.newEvent(d));
                mJoinEv.push_back(mBk.newEvent(d));
            }
        }
        if (mGraphExec) {
            cudaGraphExecDestroy(mGraphExec);
            mGraphExec = nullptr;
        }
    }

    void run()
    {
        bool graph = mOptions.cudaGraph() && mBk.getDeviceCount() == 1 && mBk.runtime() == Runtime::stream;
        for (const auto& n : mNodes) {
            /* a user lambda that writes a field with an x-face cache invalidates that cache on the HOST at every launch; a
             * replayed graph would skip it: such sequences are issued launch by launch */
            if (graph && (n.kind == Node::compute || n.kind == Node::halo) && !n.container.graphSafe()) {
                graph = false;
            }
        }
        if (!graph) {
            NEON_NVTX_PUSH("Skeleton");
            issue();
            NEON_NVTX_POP();
            return;
        }
        mBk.setDevice(0);
        cudaStream_t main = mBk.stream(0, 0);
        if (!mGraphExec) {
            cudaGraph_t g = nullptr;
            NEON_CUDA_CHECK(cudaStreamSynchronize(main));
            NEON_CUDA_CHECK(cudaStreamBeginCapture(main, cudaStreamCaptureModeThreadLocal));
            try {
                issue();
            } catch (...) {
                cudaStreamEndCapture(main, &g);
                if (g) {
                    cudaGraphDestroy(g);
                }
                throw;
            }
            NEON_CUDA_CHECK(cudaStreamEndCapture(main, &g));
            NEON_CUDA_CHECK(cudaGraphInstantiate(&mGraphExec, g, 0));
            cudaGraphDestroy(g);
        }
        NEON_CUDA_CHECK(cudaGraphLaunch(mGraphExec, main));
    }

    const std::vector<Node>& nodes() const { return mNodes; }
    const std::string&       getName() const { return mName; }

    /* the schedule in host issue order, one line per node: "stream kind name view" */
    std::string scheduleToString() const
    {
        std::string o;
        for (const auto& n : mNodes) {
            static const char* kinds[] = {"fork", "join", "halo", "compute"};
            o += std::to_string(n.stream) + " " + kinds[n.kind] + " " + n.name + " " +
                 (n.kind == Node::compute ? DataViewUtil::toString(n.view) : "-") + "\n";
        }
        return o;
    }
    /* Skeleton::ioToDot: the scheduled graph, one cluster per stream */
    void ioToDot(const std::string& fname, const std::string& graphName = "", bool = false) const
    {
        std::ofstream out(fname + ".dot");
        out << "digraph \"" << (graphName.empty() ? mName : graphName) << "\" {\n";
        int prev[8] = {-1, -1, -1, -1, -1, -1, -1, -1}, forkNode = -1;
        for (int i = 0; i < int(mNodes.size()); ++i) {
            const auto& n = mNodes[i];
            out << "  n" << i << " [label=\"" << n.name << (n.kind == Node::compute ? std::string("\\n") + DataViewUtil::toString(n.view) : "")
                << "\\nstream " << n.stream << "\"];\n";
            if (n.kind == Node::fork) {
                forkNode = i;
            }
            if (prev[n.stream] >= 0) {
                out << "  n" << prev[n.stream] << " -> n" << i << ";\n";
            } else if (forkNode >= 0 && n.stream > 0) {
                out << "  n" << forkNode << " -> n" << i << ";\n";
            }
            if (n.kind == Node::join) {
                for (int s = 1; s < 8; ++s) {
                    if (prev[s] >= 0) {
                        out << "  n" << prev[s] << " -> n" << i << ";\n";
                        prev[s] = -1;
                    }
                }
            }
            prev[n.stream] = i;
        }
        out << "}\n";
    }

   private:
    void issue()
    {
        const bool cuda = mBk.runtime() == Runtime::stream;
        const int  nDev = mBk.getDeviceCount();
        for (const auto& n : mNodes) {
            switch (n.kind) {
                case Node::fork:
                    for (int d = 0; cuda && d < nDev; ++d) {
                        mBk.setDevice(d);
                        NEON_CUDA_CHECK(cudaEventRecord(mForkEv[d], mBk.stream(d, 0)));
                        NEON_CUDA_CHECK(cudaStreamWaitEvent(mBk.stream(d, 1), mForkEv[d], 0));
                    }
                    break;
                case Node::join:
                    for (int d = 0; cuda && d < nDev; ++d) {
                        mBk.setDevice(d);
                        NEON_CUDA_CHECK(cudaEventRecord(mJoinEv[d], mBk.stream(d, 1)));
                        NEON_CUDA_CHECK(cudaStreamWaitEvent(mBk.stream(d, 0), mJoinEv[d], 0));
                    }
                    break;
                default:
                    NEON_NVTX_PUSH(n.name.c_str());
                    n.container.run(n.stream, n.view);
                    NEON_NVTX_POP();
            }
        }
    }

    Backend                  mBk;
    bool                     mHasBk = false;
    std::string              mName;
    Options                  mOptions;
    std::vector<Node>        mNodes;
    std::vector<cudaEvent_t> mForkEv, mJoinEv;
    cudaGraphExec_t          mGraphExec = nullptr;
};

}  // namespace Neon::skeleton
