// Skeleton.h — schedules a sequence of Containers on the Backend's stream sets, inserts the halo updates the declared
// stencil reads need and overlaps them with computation (OCC).
//
// Mirrors libNeonSkeleton: Skeleton::sequence / run (include/Neon/skeleton/Skeleton.h:32-65), Options(Occ, TransferMode),
// Occ (Occ.h:8-14), and what MultiXpuGraph does for a sequence (src/skeleton/internal/multiGpuGraph.cpp): dependencies
// from the Loader tokens (:43-70), OCC split of a stencil node into INTERNAL + BOUNDARY clones (:120-301), a halo-update
// node on every stencil-read edge whose consumer is not INTERNAL (:304-352), stream mapping and event insertion
// (libNeonSet/src/set/container/Graph.cpp:690-838), sequential host issue (:992-1030).
//
// The sequence becomes a dependency graph exactly as in neon_b200/skeleton.py (same algorithm, same tests):
//   parse           RAW / WAR / WAW between containers from their tokens (DependencyAnalyser), transitively reduced
//   optimizations   Occ::standard: every stencil container is split into a BOUNDARY and an INTERNAL half;
//                   Occ::extended: so are the map containers right in front of it when ALL its predecessors are maps;
//                   Occ::twoWayExtended: and the maps right behind it when predecessors and successors qualify
//   communications  a halo update in front of every stencil read whose ghost planes are not current; only halves that are
//                   not INTERNAL wait for it
//   scheduling      levels, greedy stream mapping (stream 0 = main lane for INTERNAL / STANDARD compute nodes, streams 1..
//                   are created with high priority and carry halo updates and BOUNDARY halves), one event per
//                   cross-stream edge, issue order = topological order that prefers the high-priority streams
// Differences by design: dependencies between the halves of split nodes follow the CELLS each half touches (a MAP access
// of an INTERNAL half never meets the BOUNDARY half of its producer, a STENCIL access meets both), so with Occ::extended the
// halo update waits for the BOUNDARY half of the producer only; ordering is by CUDA events only (the reference's halo
// update blocks the host on every device, SynchronizationContainer.h:37-42); the BOUNDARY view is z_local in {0, nz-1} (the
// reference folds it onto {0,1}); with one device a whole run() can be captured once into a CUDA graph and
// replayed (the launches are then frozen: call sequence() again after anything the containers look up at launch time
// changed, e.g. a field's x-face cache).  Reductions are outside the LBM path and are not modelled.
#pragma once

#include <algorithm>
#include <fstream>
#include <map>
#include <set>
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
        int            uid = -1;
        int            op = -1;    /* index of the container in the sequence (halo nodes: of the consumer) */
        std::set<int>  preds;      /* uids of the nodes this one runs after (transitively reduced) */
        int            level = 0;
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
        mEv = std::move(o.mEv);
        mWaits = std::move(o.mWaits);
        mSignals = std::move(o.mSignals);
        mForkRoots = std::move(o.mForkRoots);
        mJoins = std::move(o.mJoins);
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
        const int  nOps = int(operations.size());
        auto       isCompute = [&](int i) { return operations[i].getKind() == set::Container::Kind::compute; };
        auto       stencilReads = [&](int i) {
            std::vector<const set::Token*> r;
            for (const auto& t : operations[i].getTokens()) {
                if (t.access == set::Access::read && t.pattern == Pattern::STENCIL) {
                    r.push_back(&t);
                }
            }
            return r;
        };
        std::vector<bool> isStencil(nOps);
        for (int i = 0; i < nOps; ++i) {
            isStencil[i] = !stencilReads(i).empty();
        }

        /* ---- data dependencies between containers: one record per field uid (last writer, readers since) */
        struct Why
        {
            char    kind; /* 'R' RAW, 'A' WAR, 'W' WAW */
            Pattern early, late;
        };
        std::vector<std::map<int, std::vector<Why>>>          deps(nOps);
        std::map<uint64_t, std::pair<int, Pattern>>              lastWrite;
        std::map<uint64_t, std::vector<std::pair<int, Pattern>>> readers;
        for (int i = 0; i < nOps; ++i) {
            const auto& tokens = operations[i].getTokens();
            for (const auto& t : tokens) {
                if (t.access == set::Access::read && lastWrite.count(t.uid) && lastWrite[t.uid].first != i) {
                    deps[i][lastWrite[t.uid].first].push_back({'R', lastWrite[t.uid].second, t.pattern});
                }
            }
            for (const auto& t : tokens) {
                if (t.access != set::Access::write) {
                    continue;
                }
                for (const auto& r : readers[t.uid]) {
                    if (r.first != i) {
                        deps[i][r.first].push_back({'A', r.second, t.pattern});
                    }
                }
                if (lastWrite.count(t.uid) && lastWrite[t.uid].first != i) {
                    deps[i][lastWrite[t.uid].first].push_back({'W', lastWrite[t.uid].second, t.pattern});
                }
            }
            for (const auto& t : tokens) {
                if (t.access == set::Access::read) {
                    readers[t.uid].push_back({i, t.pattern});
                }
            }
            for (const auto& t : tokens) {
                if (t.access == set::Access::write) {
                    lastWrite[t.uid] = {i, t.pattern};
                    auto& r = readers[t.uid];
                    r.erase(std::remove_if(r.begin(), r.end(), [&](const std::pair<int, Pattern>& e) { return e.first != i; }), r.end());
                }
            }
        }
        /* direct predecessors / successors (Graph::removeRedundantDependencies) */
        std::vector<std::set<int>> reach(nOps), direct(nOps), succ(nOps);
        for (int i = 0; i < nOps; ++i) {
            for (auto it = deps[i].rbegin(); it != deps[i].rend(); ++it) {
                const int j = it->first;
                if (!reach[i].count(j)) {
                    direct[i].insert(j);
                }
                reach[i].insert(j);
                reach[i].insert(reach[j].begin(), reach[j].end());
            }
            for (int j : direct[i]) {
                succ[j].insert(i);
            }
        }

        /* ---- OCC: which containers are split (multiGpuGraph.cpp:120-301) */
        std::set<int> split;
        if (multi && options.occ() != Occ::none) {
            for (int st = 0; st < nOps; ++st) {
                if (!isStencil[st] || !isCompute(st)) {
                    continue;
                }
                split.insert(st);
                if (options.occ() == Occ::standard) {
                    continue;
                }
                auto mapsOnly = [&](const std::set<int>& around) {
                    if (around.empty()) {
                        return false;
                    }
                    for (int j : around) {
                        if (!isCompute(j) || isStencil[j]) {
                            return false;
                        }
                    }
                    return true;
                };
                const bool beforeOk = mapsOnly(direct[st]), afterOk = mapsOnly(succ[st]);
                if (options.occ() == Occ::extended && beforeOk) {
                    split.insert(direct[st].begin(), direct[st].end());
                }
                if (options.occ() == Occ::twoWayExtended && beforeOk && afterOk) {
                    split.insert(direct[st].begin(), direct[st].end());
                    split.insert(succ[st].begin(), succ[st].end());
                }
            }
        }

        /* ---- nodes */
        std::vector<Node> nodes;
        auto              add = [&](Node::Kind kind, const std::string& nm, DataView view, const set::Container& c, int op) {
            Node n;
            n.kind = kind;
            n.name = nm;
            n.view = view;
            n.container = c;
            n.uid = int(nodes.size());
            n.op = op;
            nodes.push_back(n);
            return n.uid;
        };
        auto overlap = [](DataView a, DataView b) { return a == DataView::STANDARD || b == DataView::STANDARD || a == b; };
        std::vector<std::vector<int>>    pieces(nOps), haloIn(nOps);
        std::map<uint64_t, bool>             fresh;        /* field uid -> ghost planes current within this sequence */
        std::map<uint64_t, std::vector<int>> ghostReaders; /* field uid -> nodes that read the ghost planes since the last update */
        for (int i = 0; i < nOps; ++i) {
            const auto& c = operations[i];
            if (multi && isCompute(i)) {
                for (const set::Token* t : stencilReads(i)) {
                    if (fresh[t->uid] || !t->newHaloUpdate) {
                        continue;
                    }
                    auto      h = t->newHaloUpdate(t->semantic, options.transferMode(), Execution::device);
                    const int hu = add(Node::halo, h.getName(), DataView::STANDARD, h, i);
                    haloIn[i].push_back(hu);
                    /* the update overwrites ghost planes: after every earlier reader of them (WAR) and after the faces it sends
                     * exist (RAW: the non-INTERNAL halves of the last writer) */
                    nodes[hu].preds.insert(ghostReaders[t->uid].begin(), ghostReaders[t->uid].end());
                    for (int j = i - 1; j >= 0; --j) {
                        bool writes = false;
                        for (const auto& w : operations[j].getTokens()) {
                            writes = writes || (w.access == set::Access::write && w.uid == t->uid);
                        }
                        if (writes) {
                            for (int pj : pieces[j]) {
                                if (nodes[pj].view != DataView::INTERNAL) {
                                    nodes[hu].preds.insert(pj);
                                }
                            }
                            break;
                        }
                    }
                    fresh[t->uid] = true;
                    ghostReaders[t->uid].clear();
                }
            }
            const Node::Kind kind = c.getKind() == set::Container::Kind::halo ? Node::halo : Node::compute;
            if (split.count(i)) {
                pieces[i] = {add(kind, c.getName(), DataView::BOUNDARY, c, i), add(kind, c.getName(), DataView::INTERNAL, c, i)};
            } else {
                pieces[i] = {add(kind, c.getName(), DataView::STANDARD, c, i)};
            }
            for (const auto& dj : deps[i]) {
                for (int pi : pieces[i]) {
                    for (int pj : pieces[dj.first]) {
                        for (const Why& w : dj.second) {
                            const bool stencil = (w.kind == 'R' ? w.late : w.early) == Pattern::STENCIL;
                            if (stencil || overlap(nodes[pi].view, nodes[pj].view)) {
                                nodes[pi].preds.insert(pj);
                                break;
                            }
                        }
                    }
                }
            }
            for (int pi : pieces[i]) {
                if (nodes[pi].view != DataView::INTERNAL) {
                    nodes[pi].preds.insert(haloIn[i].begin(), haloIn[i].end());
                    for (const set::Token* t : stencilReads(i)) {
                        ghostReaders[t->uid].push_back(pi);
                    }
                }
            }
            for (const auto& t : c.getTokens()) {
                if (t.access == set::Access::write) {
                    fresh[t.uid] = false;
                }
            }
        }
        schedule(nodes);

        int width = 1;
        for (const auto& n : mNodes) {
            width = std::max(width, n.stream + 1);
        }
        mBk.setAvailableStreamSet(width);
        if (mBk.runtime() == Runtime::stream) { /* every event exists before the first run: an iteration allocates nothing */
            const int nDev = mBk.getDeviceCount();
            if (mForkEv.empty()) {
                for (int d = 0; d < nDev; ++d) {
                    mForkEv.push_back(mBk.newEvent(d));
                }
            }
            for (int uid : mSignals) {
                auto& ev = mEv[uid];
                while (int(ev.size()) < nDev) {
                    ev.push_back(mBk.newEvent(int(ev.size())));
                }
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

    /* the schedule in host issue order, one line per node: "stream kind name view"; bracketed by the fork / join of the
     * side streams when more than one stream is used */
    std::string scheduleToString() const
    {
        std::string o;
        bool        side = false;
        for (const auto& n : mNodes) {
            side = side || n.stream != 0;
        }
        if (side) {
            o += "0 fork fork -\n";
        }
        for (const auto& n : mNodes) {
            o += std::to_string(n.stream) + " " + (n.kind == Node::halo ? "halo" : "compute") + " " + n.name + " " +
                 (n.kind == Node::compute ? DataViewUtil::toString(n.view) : "-") + "\n";
        }
        if (side) {
            o += "0 join join -\n";
        }
        return o;
    }
    /* "kind name view <- kind name view, ..." for every node: the scheduled dependency graph */
    std::string dependenciesToString() const
    {
        auto key = [](const Node& n) {
            return std::string(n.kind == Node::halo ? "halo " : "compute ") + n.name + " " + (n.kind == Node::compute ? DataViewUtil::toString(n.view) : "-");
        };
        std::map<int, const Node*> byUid;
        for (const auto& n : mNodes) {
            byUid[n.uid] = &n;
        }
        std::string o;
        for (const auto& n : mNodes) {
            std::vector<std::string> p;
            for (int u : n.preds) {
                p.push_back(key(*byUid.at(u)));
            }
            std::sort(p.begin(), p.end());
            o += key(n) + " <-";
            for (const auto& e : p) {
                o += " [" + e + "]";
            }
            o += "\n";
        }
        return o;
    }
    /* Skeleton::ioToDot: the scheduled graph */
    void ioToDot(const std::string& fname, const std::string& graphName = "", bool = false) const
    {
        std::ofstream out(fname + ".dot");
        out << "digraph \"" << (graphName.empty() ? mName : graphName) << "\" {\n";
        for (const auto& n : mNodes) {
            out << "  n" << n.uid << " [label=\"" << n.name << (n.kind == Node::compute ? std::string("\\n") + DataViewUtil::toString(n.view) : "")
                << "\\nstream " << n.stream << "\"];\n";
            for (int p : n.preds) {
                out << "  n" << p << " -> n" << n.uid << ";\n";
            }
        }
        out << "}\n";
    }

   private:
    /* levels, streams, events, issue order (libNeonSet/src/set/container/Graph.cpp:652-661, 690-838) */
    void schedule(std::vector<Node>& nodes)
    {
        const int                  n = int(nodes.size());
        std::vector<std::set<int>> reach(n);
        for (auto& nd : nodes) { /* uids are a topological order: every predecessor was created earlier */
            std::set<int> kept;
            for (auto it = nd.preds.rbegin(); it != nd.preds.rend(); ++it) {
                if (!reach[nd.uid].count(*it)) {
                    kept.insert(*it);
                }
                reach[nd.uid].insert(*it);
                reach[nd.uid].insert(reach[*it].begin(), reach[*it].end());
            }
            nd.preds = kept;
            nd.level = 0;
            for (int p : nd.preds) {
                nd.level = std::max(nd.level, nodes[p].level + 1);
            }
        }
        auto mainLane = [](const Node& nd) { return nd.kind == Node::compute && nd.view != DataView::BOUNDARY; };
        int  levels = 0;
        for (const auto& nd : nodes) {
            levels = std::max(levels, nd.level + 1);
        }
        for (int lvl = 0; lvl < levels; ++lvl) {
            std::vector<int> todo;
            for (const auto& nd : nodes) {
                if (nd.level == lvl) {
                    todo.push_back(nd.uid);
                }
            }
            std::stable_sort(todo.begin(), todo.end(), [&](int a, int b) { return mainLane(nodes[a]) && !mainLane(nodes[b]); });
            std::set<int> booked;
            for (int u : todo) {
                Node&            nd = nodes[u];
                const bool       lane0 = mainLane(nd) || todo.size() == 1;
                std::vector<int> mine;
                for (int p : nd.preds) {
                    mine.push_back(nodes[p].stream);
                }
                std::stable_sort(mine.begin(), mine.end(), [&](int a, int b) {
                    const bool wa = (a == 0) != mainLane(nd), wb = (b == 0) != mainLane(nd);
                    return wa != wb ? !wa : a < b;
                });
                int st = -1;
                for (int c : mine) {
                    if (!booked.count(c)) {
                        st = c;
                        break;
                    }
                }
                if (st < 0) {
                    st = lane0 ? 0 : 1;
                    while (booked.count(st)) {
                        ++st;
                    }
                }
                booked.insert(st);
                nd.stream = st;
            }
        }
        /* issue order: topological, high-priority streams first, then lower levels */
        std::vector<Node> order;
        std::set<int>     done;
        std::vector<bool> placed(n, false);
        while (int(order.size()) < n) {
            int best = -1;
            for (int u = 0; u < n; ++u) {
                if (placed[u] || !std::includes(done.begin(), done.end(), nodes[u].preds.begin(), nodes[u].preds.end())) {
                    continue;
                }
                if (best < 0) {
                    best = u;
                    continue;
                }
                const auto ka = std::make_tuple(nodes[u].stream == 0, nodes[u].level, u);
                const auto kb = std::make_tuple(nodes[best].stream == 0, nodes[best].level, best);
                if (ka < kb) {
                    best = u;
                }
            }
            placed[best] = true;
            done.insert(best);
            order.push_back(nodes[best]);
        }
        mWaits.clear();
        mSignals.clear();
        mForkRoots.clear();
        mJoins.clear();
        for (const auto& nd : nodes) {
            bool sameStreamPred = false;
            for (int p : nd.preds) {
                if (nodes[p].stream != nd.stream) {
                    mWaits[nd.uid].push_back(p);
                    mSignals.insert(p);
                } else {
                    sameStreamPred = true;
                }
            }
            if (nd.stream != 0 && !sameStreamPred && mWaits[nd.uid].empty()) {
                mForkRoots.insert(nd.uid);
            }
        }
        std::map<int, int> lastOn;
        for (const auto& nd : order) {
            lastOn[nd.stream] = nd.uid;
        }
        for (const auto& e : lastOn) {
            if (e.first != 0) {
                mJoins.push_back(e.second);
                mSignals.insert(e.second);
            }
        }
        mNodes = std::move(order);
    }

    void issue()
    {
        const bool cuda = mBk.runtime() == Runtime::stream;
        const int  nDev = mBk.getDeviceCount();
        if (cuda && !mForkRoots.empty()) {
            for (int d = 0; d < nDev; ++d) {
                mBk.setDevice(d);
                NEON_CUDA_CHECK(cudaEventRecord(mForkEv[d], mBk.stream(d, 0)));
            }
        }
        for (const auto& n : mNodes) {
            if (cuda) {
                for (int d = 0; d < nDev; ++d) {
                    mBk.setDevice(d);
                    if (mForkRoots.count(n.uid)) {
                        NEON_CUDA_CHECK(cudaStreamWaitEvent(mBk.stream(d, n.stream), mForkEv[d], 0));
                    }
                    auto w = mWaits.find(n.uid);
                    if (w != mWaits.end()) {
                        for (int p : w->second) {
                            NEON_CUDA_CHECK(cudaStreamWaitEvent(mBk.stream(d, n.stream), mEv.at(p)[d], 0));
                        }
                    }
                }
            }
            NEON_NVTX_PUSH(n.name.c_str());
            n.container.run(n.stream, n.view);
            NEON_NVTX_POP();
            if (cuda && mSignals.count(n.uid)) {
                for (int d = 0; d < nDev; ++d) {
                    mBk.setDevice(d);
                    NEON_CUDA_CHECK(cudaEventRecord(mEv.at(n.uid)[d], mBk.stream(d, n.stream)));
                }
            }
        }
        if (cuda) {
            for (int uid : mJoins) {
                for (int d = 0; d < nDev; ++d) {
                    mBk.setDevice(d);
                    NEON_CUDA_CHECK(cudaStreamWaitEvent(mBk.stream(d, 0), mEv.at(uid)[d], 0));
                }
            }
        }
    }

    Backend                  mBk;
    bool                     mHasBk = false;
    std::string              mName;
    Options                  mOptions;
    std::vector<Node>        mNodes;
    std::vector<cudaEvent_t>                mForkEv;    /* per device */
    std::map<int, std::vector<cudaEvent_t>> mEv;        /* node uid -> per device */
    std::map<int, std::vector<int>>         mWaits;     /* node uid -> uids on other streams it waits for */
    std::set<int>                           mSignals;   /* uids that record an event */
    std::set<int>                           mForkRoots; /* side-stream nodes without a predecessor: wait for the fork event */
    std::vector<int>                        mJoins;     /* last node of every side stream */
    cudaGraphExec_t          mGraphExec = nullptr;
};

}  // namespace Neon::skeleton
