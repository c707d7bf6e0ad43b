// host-logic — checks of the C++ veneer that need no GPU (Runtime::openmp backend: host mirrors and scheduling only).
//
// Partition rule (dGrid_imp.h:43-62), spans per data view (dSpan_imp.h:6-43 with the corrected BOUNDARY map), the schedule a
// Skeleton builds for Occ::none / Occ::standard (multiGpuGraph.cpp:120-143,304-352), Loader tokens, the CellType <-> flag-word
// codec (CellType.h:33-34 vs include/neon_lbm.h), bGrid blocks / layers / ghost blocks (bGrid_imp.h:7-185), the report
// writer, and that compute containers refuse to run without Runtime::stream (no CPU fallback).
#include <cstdio>
#include <fstream>
#include <set>
#include <string>

#include "Neon/Neon.h"
#include "Neon/Report.h"
#include "Neon/domain/bGrid.h"
#include "Neon/domain/dGrid.h"
#include "Neon/lbm/Lbm.h"
#include "Neon/skeleton/Skeleton.h"

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);    \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

int main()
{
    using namespace Neon;
    init();
    D3Q19Template<float, float> lattice((Backend()));
    CHECK(lattice.c_vect.size() == 19 && lattice.c_vect[9] == index_3d(0, 0, 0));
    for (int k = 0; k < 19; ++k) {
        CHECK(lattice.c_vect[lattice.opp_vect[k]] == -lattice.c_vect[k]);
    }
    double wsum = 0;
    for (double w : lattice.t_vect) {
        wsum += w;
    }
    CHECK(std::abs(wsum - 1.0) < 1e-15);
    D3Q27Template<double, double> l27((Backend()));
    CHECK(l27.c_vect.size() == 27 && l27.c_vect[0] == index_3d(0, 0, 0) && l27.c_vect[1] == index_3d(0, 0, -1) &&
          l27.c_vect[26] == index_3d(1, 1, 1) && l27.opp_vect[9] == 18 && l27.opp_vect[13] == 26);

    // ---- partitioning: 10 planes over 3 devices -> 4, 3, 3 with origins 0, 4, 7; one ghost plane per side
    Backend bk({0, 0, 0}, Runtime::openmp);
    dGrid   grid(bk, {20, 6, 10}, [](const index_3d&) { return true; }, lattice.c_vect);
    CHECK(grid.getNumPartitions() == 3 && grid.zHalo() == 1 && grid.latticeQ() == 19);
    CHECK(bk.devSet().setCardinality() == 3 && bk.devSet().devId(2) == 0);
    CHECK(grid.nzLocal(0) == 4 && grid.nzLocal(1) == 3 && grid.nzLocal(2) == 3);
    CHECK(grid.zOrigin(0) == 0 && grid.zOrigin(1) == 4 && grid.zOrigin(2) == 7);
    const nlbm_dense_desc d1 = grid.descOf(1);
    CHECK(d1.pitch_y == 128 && d1.pitch_z == 128 * 6 && d1.pitch_q == 128 * 6 * 5 && d1.z_origin == 4 && d1.gnz == 10);

    // ---- spans: STANDARD all planes, INTERNAL [1, nz-1), BOUNDARY {0, nz-1}
    {
        dSpan st = grid.getSpan(0, DataView::STANDARD), in = grid.getSpan(0, DataView::INTERNAL), bd = grid.getSpan(0, DataView::BOUNDARY);
        CHECK(st.nzView == 4 && in.nzView == 2 && bd.nzView == 2);
        dIdx     i;
        std::set<int> seen;
        for (int z = 0; z < in.nzView; ++z) {
            CHECK(in.setAndValidate(i, 3, 2, z));
            seen.insert(i.z);
        }
        for (int z = 0; z < bd.nzView; ++z) {
            CHECK(bd.setAndValidate(i, 3, 2, z));
            CHECK(seen.insert(i.z).second);  // the two views are disjoint
        }
        CHECK(seen == std::set<int>({0, 1, 2, 3}));  // and tile the partition
        CHECK(!st.setAndValidate(i, 20, 0, 0) && !st.setAndValidate(i, 0, 6, 0) && !st.setAndValidate(i, 0, 0, 4));
    }
    {
        Backend one({0}, Runtime::openmp);
        dGrid   g1(one, {8, 8, 8}, [](const index_3d&) { return true; }, lattice.c_vect);
        CHECK(g1.zHalo() == 0 && g1.getSpan(0, DataView::INTERNAL).nzView == 8 && g1.getSpan(0, DataView::BOUNDARY).nzView == 0);
    }

    // ---- fields: host mirror, neighbour validity, codec
    auto pop0 = grid.newField<float, 19>("pop0", 19, 0.f);
    auto pop1 = grid.newField<float, 19>("pop1", 19, 0.f);
    auto flag = grid.newField<CellType, 1>("flag", 1, CellType());
    pop0.forEachActiveCell([](const index_3d& p, const int& k, float& v) { v = float(p.x + 100 * p.y + 10000 * p.z) + 0.01f * k; });
    CHECK(pop0(index_3d(3, 2, 7), 5) == float(3 + 200 + 70000) + 0.05f);
    CHECK(pop0(index_3d(-1, 0, 0), 0) == 0.f);  // outside value
    {
        const auto& part = pop0.getPartition(1);  // planes 4..6
        dIdx        i{0, 0, 0};
        CHECK(part.isNghValid(i, 0, 0, -1) && !part.isNghValid(i, -1, 0, 0) && !part.isNghValid(i, 0, -1, 0));
        CHECK(part.getGlobalIndex(i) == index_3d(0, 0, 4));
        const auto& last = pop0.getPartition(2);
        dIdx        top{5, 5, 2};  // global z = 9
        CHECK(!last.isNghValid(top, 0, 0, 1) && last.isNghValid(top, 1, 0, -1));
        CHECK(part.offset(1, 2, 0, 3) == 3 * d1.pitch_q + 1 * d1.pitch_z + 2 * d1.pitch_y + 1);
    }
    {
        CellType       c(CellType::movingWall, 0x0012345u);
        const uint32_t w = domain::FlagWordCodec<CellType>::pack(c);
        CHECK(NLBM_FLAG_CLASS(w) == NLBM_MOVING_WALL && (w & NLBM_FLAG_MASK_BITS) == 0x0012345u);
        const CellType back = domain::FlagWordCodec<CellType>::unpack(w);
        CHECK(back.classification == CellType::movingWall && back.wallNghBitflag == 0x0012345u);
        CHECK(CellType(CellType::undefined).classification == CellType::undefined);
        CHECK(CellType(7).classification == CellType::bulk);  // the reference's int constructor ignores its argument
    }

    // ---- containers, tokens, schedule
    using Pop = dGrid::Field<float, 19>;
    using Tools = LbmContainers<D3Q19Template<float, float>, Pop, float>;
    auto it = Tools::iteration(set::StencilSemantic::streaming, pop0, flag, 1.2f, pop1);
    CHECK(it.getTokens().size() == 3);
    CHECK(it.getTokens()[0].access == set::Access::read && it.getTokens()[0].pattern == Pattern::STENCIL &&
          it.getTokens()[0].semantic == set::StencilSemantic::streaming && bool(it.getTokens()[0].newHaloUpdate));
    CHECK(it.getTokens()[1].access == set::Access::write && it.getTokens()[2].access == set::Access::read);
    bool refused = false;
    try {
        it.run(0);
    } catch (const NeonException& e) {
        refused = std::string(e.what()).find("no CPU fallback") != std::string::npos;
    }
    CHECK(refused);
    bool aliasRefused = false;
    try {
        Tools::iteration(set::StencilSemantic::streaming, pop0, flag, 1.2f, pop0);
    } catch (const NeonException&) {
        aliasRefused = true;
    }
    CHECK(aliasRefused);
    {
        skeleton::Skeleton sk(bk);
        sk.sequence({it}, "none", skeleton::Options(skeleton::Occ::none, set::TransferMode::get));
        CHECK(sk.scheduleToString() ==
              "0 halo haloUpdate(pop0,streaming,get) -\n"
              "0 compute LBM_iteration_D3Q19 STANDARD\n");
        sk.sequence({it}, "occ", skeleton::Options(skeleton::Occ::standard, set::TransferMode::put));
        CHECK(sk.scheduleToString() ==
              "0 fork fork -\n"
              "1 halo haloUpdate(pop0,streaming,put) -\n"
              "1 compute LBM_iteration_D3Q19 BOUNDARY\n"
              "0 compute LBM_iteration_D3Q19 INTERNAL\n"
              "0 join join -\n");  // the high-priority side stream is issued first
        CHECK(bk.getStreamSetCount() == 2);
        sk.ioToDot("/tmp/neon_b200_host_logic_graph");
        std::ifstream dot("/tmp/neon_b200_host_logic_graph.dot");
        std::string   all((std::istreambuf_iterator<char>(dot)), std::istreambuf_iterator<char>());
        CHECK(all.find("INTERNAL") != std::string::npos && all.find("BOUNDARY") != std::string::npos);
        Backend            one({0}, Runtime::openmp);
        dGrid              g1(one, {8, 8, 8}, [](const index_3d&) { return true; }, lattice.c_vect);
        auto               a = g1.newField<float, 19>("a", 19, 0.f), b = g1.newField<float, 19>("b", 19, 0.f);
        auto               f = g1.newField<CellType, 1>("f", 1, CellType());
        skeleton::Skeleton s1(one);
        s1.sequence({Tools::iteration(set::StencilSemantic::streaming, a, f, 1.f, b)}, "single", skeleton::Options(skeleton::Occ::standard, set::TransferMode::get));
        CHECK(s1.scheduleToString() == "0 compute LBM_iteration_D3Q19 STANDARD\n");  // one device: no halo, no split
    }
    {
        // ---- the Skeleton's graph for map -> stencil -> map under every Occ (multiGpuGraph.cpp:120-301): containers whose
        // body is never run, only their tokens matter here
        auto fa = grid.newField<float, 19>("fa", 19, 0.f), fb = grid.newField<float, 19>("fb", 19, 0.f);
        auto fc = grid.newField<float, 19>("fc", 19, 0.f), fd = grid.newField<float, 19>("fd", 19, 0.f);
        auto mk = [&](const std::string& name, const Pop& in, Pop& out, bool stencil) {
            return set::Container::factoryDeviceManaged(name, bk, [&](SetIdx, set::Loader& L) {
                if (stencil) {
                    L.load(in, Pattern::STENCIL, set::StencilSemantic::standard);
                } else {
                    L.load(in);
                }
                L.load(out);
                return [](int, DataView) {};
            });
        };
        std::vector<set::Container> seq = {mk("M1", fa, fb, false), mk("S", fb, fc, true), mk("M2", fc, fd, false)};
        skeleton::Skeleton          sk(bk);
        sk.sequence(seq, "std", skeleton::Options(skeleton::Occ::standard, set::TransferMode::get));
        CHECK(sk.dependenciesToString() ==
              "compute M1 STANDARD <-\n"
              "halo haloUpdate(fb,grid,get) - <- [compute M1 STANDARD]\n"
              "compute S BOUNDARY <- [halo haloUpdate(fb,grid,get) -]\n"
              "compute S INTERNAL <- [compute M1 STANDARD]\n"
              "compute M2 STANDARD <- [compute S BOUNDARY] [compute S INTERNAL]\n");
        sk.sequence(seq, "ext", skeleton::Options(skeleton::Occ::extended, set::TransferMode::get));
        const std::string ext = sk.dependenciesToString();
        CHECK(ext.find("halo haloUpdate(fb,grid,get) - <- [compute M1 BOUNDARY]\n") != std::string::npos);  // overlaps M1's INTERNAL half
        CHECK(ext.find("compute S INTERNAL <- [compute M1 BOUNDARY] [compute M1 INTERNAL]\n") != std::string::npos);
        CHECK(ext.find("compute M2 STANDARD") != std::string::npos);
        CHECK(sk.scheduleToString().find("compute M1 BOUNDARY") < sk.scheduleToString().find("compute M1 INTERNAL"));
        sk.sequence(seq, "two", skeleton::Options(skeleton::Occ::twoWayExtended, set::TransferMode::get));
        const std::string two = sk.dependenciesToString();
        CHECK(two.find("compute M2 INTERNAL <- [compute S INTERNAL]\n") != std::string::npos);
        CHECK(two.find("compute M2 BOUNDARY <- [compute S BOUNDARY]\n") != std::string::npos);
        sk.sequence(seq, "none", skeleton::Options(skeleton::Occ::none, set::TransferMode::get));
        CHECK(sk.scheduleToString() ==
              "0 compute M1 STANDARD\n"
              "0 halo haloUpdate(fb,grid,get) -\n"
              "0 compute S STANDARD\n"
              "0 compute M2 STANDARD\n");
    }
    {
        auto halo = pop0.newHaloUpdate(set::StencilSemantic::streaming, set::TransferMode::get);
        auto impl = halo.as<Neon::detail::DenseHaloImpl>();
        CHECK(impl && impl->latticeQ == 19 && impl->bytesPerFace() == size_t(5) * 128 * 6 * 4);
        auto grid19 = pop0.newHaloUpdate(set::StencilSemantic::standard, set::TransferMode::put).as<Neon::detail::DenseHaloImpl>();
        CHECK(grid19->latticeQ == 0 && grid19->bytesPerFace() == size_t(19) * 128 * 6 * 4);
    }

    // ---- bGrid: 20 x 12 x 40 cells = 3 x 2 x 5 blocks; a hole deactivates one whole block and part of another
    {
        Backend two({0, 0}, Runtime::openmp);
        auto    active = [](const index_3d& p) { return !(p.x < 8 && p.y < 8 && p.z < 8) && !(p.x >= 16 && p.y == 11 && p.z == 39); };
        bGrid   bg(two, {20, 12, 40}, active, lattice.c_vect);
        CHECK(bg.getNumPartitions() == 2 && bg.getNumBlocks() == 29 && bg.latticeQ() == 19);
        CHECK(bg.getNumActiveCells() == size_t(20) * 12 * 40 - 512 - 4);
        const auto& p0 = bg.partition(0);  // layers 0..2 (3 of 5), p1: layers 3..4
        const auto& p1 = bg.partition(1);
        CHECK(p0.nBlocks == 17 && p0.nDown == 0 && p0.nUp == 6 && p0.nGhostDown == 0 && p0.nGhostUp == 6 && p0.nAlloc == 23);
        CHECK(p1.nBlocks == 12 && p1.nDown == 6 && p1.nUp == 0 && p1.nGhostDown == 6 && p1.nGhostUp == 0 && p1.nAlloc == 18);
        CHECK(p0.coords[0][0] == 0 && p0.coords[0][1] == 0 && p0.coords[0][2] == 1);  // block (0,0,0) is the hole
        CHECK(!bg.isActive(index_3d(3, 3, 3)) && bg.isActive(index_3d(8, 0, 0)) && !bg.isActive(index_3d(17, 11, 39)));
        auto bf = bg.newField<float, 19>("bpop", 19, -1.f);
        CHECK(bf(index_3d(3, 3, 3), 0) == -1.f);
        size_t visited = 0;
        bf.forEachActiveCell([&](const index_3d&, const int& k, float&) { visited += k == 0; }, computeMode_t::seq);
        CHECK(visited == bg.getNumActiveCells());
    }

    // ---- report
    {
        Report r("host-logic");
        r.addMember("N", 64);
        r.addMember("omega", 1.25);
        r.addMember("devices", std::vector<int>{0, 1});
        r.addMember("grid", std::string("dGrid"));
        r.addMember("benchmark", true);
        r.addMember("MLUPS", std::vector<double>{1.5, 2.5});
        const std::string s = r.dump();
        CHECK(s.find("\"Record Name\": \"host-logic\"") != std::string::npos && s.find("\"devices\": [0, 1]") != std::string::npos &&
              s.find("\"benchmark\": true") != std::string::npos && s.find("\"MLUPS\": [1.5, 2.5]") != std::string::npos);
    }
    std::printf(failures ? "host-logic: %d check(s) FAILED\n" : "host-logic: all checks passed\n", failures);
    return failures ? 1 : 0;
}
