// lbm_step.cuh — the fused pull-stream + BGK collide kernel for dense (dGrid) partitions.
//
// Replaces the reference's generic lambda kernel
//   denseSpan::launchLambdaOnSpanCUDA  (libNeonSet/include/Neon/set/LambdaExecutor.h:12-39)
// carrying LbmContainers::iteration   (benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h:285-325)
// = pullStream (:99-168) + macroscopic (:172-195) + collideBgkUnrolled (:199-282), and for D3Q27
// apps/lbmMultiRes/{stream.h:5-49, collide.h:286-354, util.h:47-62}.
//
// Design (B200, HBM-bound: every population element is read once and written once per iteration):
//  * SoA planes with a 512-byte aligned row pitch; one warp owns a tile of RPW rows x (32/RPW)*VEC consecutive cells;
//    every population row is fetched with ONE aligned 16-byte load per thread.  The +-1 x shift of the
//    populations with c_x != 0 is served inside the warp by a shuffle; only the two edge lanes issue a
//    4/8-byte load, which hits a sector the neighbouring warp streams anyway.
//  * a per-row chunk summary (lbm_common.cuh) lets warps skip flag loads, wall fix-ups and predicated
//    stores where all cells are plain bulk; warps without bulk cells exit at once.
//  * wall handling (half-way bounce-back with the wall's stored population, moving lid included) is a
//    per-cell fix-up executed only by cells whose wallNghBitflag is non-zero.
//  * results leave through 16-byte streaming stores; non-bulk cells are never written (LbmTools.h:304).
#pragma once
#include <mutex>
#include <type_traits>
#include <utility>

#include "lbm_collide_exact.cuh"
#include "lbm_common.cuh"

namespace nlbm {

constexpr int kStepThreads = 256;

// =============================================================== collide policies
// FMAD is only a symbol tag: the *_ref.cu translation unit is compiled with -fmad=false, the *_fast.cu
// one with contraction on; the tag keeps their instantiations apart.

// D3Q19, expression-for-expression the reference (LbmTools.h:172-195, 199-282, 312-314): operand types
// and association are kept so that the usual arithmetic conversions round exactly as the CPU build does.
template <typename S, typename C, int FMAD>
struct CollideD3Q19Ref
{
    static constexpr int Q = 19;
    using Compute = C;
    __device__ __forceinline__ static void run(S (&p)[19], const C omega)
    {
#define P(i) ((C)p[i])
        const C X_M1 = P(0) + P(3) + P(4) + P(5) + P(6);
        const C X_P1 = P(10) + P(13) + P(14) + P(15) + P(16);
        const C X_0 = P(9) + P(1) + P(2) + P(7) + P(8) + P(11) + P(12) + P(17) + P(18);
        const C Y_M1 = P(1) + P(3) + P(7) + P(8) + P(14);
        const C Y_P1 = P(4) + P(11) + P(13) + P(17) + P(18);
        const C Z_M1 = P(2) + P(5) + P(7) + P(16) + P(18);
        const C Z_P1 = P(6) + P(8) + P(12) + P(15) + P(17);
#undef P
        const C rho = X_M1 + X_P1 + X_0;
        const C u0 = (X_P1 - X_M1) / rho;
        const C u1 = (Y_P1 - Y_M1) / rho;
        const C u2 = (Z_P1 - Z_M1) / rho;
        const C usqr = 1.5 * (u0 * u0 + u1 * u1 + u2 * u2);
        const C cu[9] = {u0, u1, u2, u0 + u1, u0 - u1, u0 + u2, u0 - u2, u1 + u2, u1 - u2};
#pragma unroll
        for (int g = 0; g < 9; ++g) {
            const double w = g < 3 ? (1. / 18.) : (1. / 36.);
            const C      eq = rho * w * (1. - 3. * cu[g] + 4.5 * cu[g] * cu[g] - usqr);
            const C      eqopp = eq + rho * w * 6. * cu[g];
            const C      o_go = (1. - omega) * (C)p[g] + omega * eq;
            const C      o_bk = (1. - omega) * (C)p[g + 10] + omega * eqopp;
            p[g] = (S)o_go;
            p[g + 10] = (S)o_bk;
        }
        const C eq9 = rho * (1. / 3.) * (1. - usqr);
        const C o9 = (1. - omega) * (C)p[9] + omega * eq9;
        p[9] = (S)o9;
    }
};

// D3Q19 fp32, the reference's bits at a fraction of the conversions (lbm_collide_exact.cuh): the fast evaluation where its
// guard holds (positive populations, |u| < 0.1), else the plain one — out of line, it is rare and would double the code.
static __device__ __noinline__ void collideD3Q19ExactSlow(float* p, const float omega)
{
    float f[19];
#pragma unroll
    for (int q = 0; q < 19; ++q)
        f[q] = p[q];
    exact::collideD3Q19<0>(f, omega);
#pragma unroll
    for (int q = 0; q < 19; ++q)
        p[q] = f[q];
}
template <int FMAD>
struct CollideD3Q19Exact
{
    static constexpr int  Q = 19;
    static constexpr bool kBulkOnly = true;  // skip cells whose result is discarded (their zeros would fail the guard)
    using Compute = float;
    __device__ __forceinline__ static void run(float (&p)[19], const float omega)
    {
        if (!exact::collideD3Q19<1>(p, omega)) {
            float t[19];
#pragma unroll
            for (int q = 0; q < 19; ++q)
                t[q] = p[q];
            collideD3Q19ExactSlow(t, omega);
#pragma unroll
            for (int q = 0; q < 19; ++q)
                p[q] = t[q];
        }
    }
};
template <class COL, typename = void>
struct BulkOnly
{
    static constexpr bool value = false;
};
template <class COL>
struct BulkOnly<COL, std::enable_if_t<COL::kBulkOnly>>
{
    static constexpr bool value = true;
};

// D3Q19, storage-precision arithmetic with fused multiply-add (agrees with the reference within the
// north-star tolerance, not bitwise).
template <typename T, int FMAD>
struct CollideD3Q19Fast
{
    static constexpr int Q = 19;
    using Compute = T;
    __device__ __forceinline__ static T rcp(T v)
    {
        if constexpr (sizeof(T) == 4)
            return __frcp_rn(v);
        else
            return T(1) / v;
    }
    __device__ __forceinline__ static T fma_(T a, T b, T c)
    {
        if constexpr (sizeof(T) == 4)
            return __fmaf_rn(a, b, c);
        else
            return __fma_rn(a, b, c);
    }
    // fp32: float arithmetic only, but with the reference's ROUNDING POINTS.  In the reference (ComputeFP = float) the moments,
    // usqr and cu are float operations — repeated here operation for operation (the three divisions through the shared
    // reciprocal of lbm_collide_exact.cuh) — eq and eqopp are double expressions rounded to float, omega*eq is a float product
    // and the relaxation a double sum rounded to float.  Here rho*w is carried as two floats, so eq = rh + (rh*s + rl) is
    // within ~0.6 ulp of the correctly rounded value; eqopp, omega*eq and the relaxation (one FMA: both products exact)
    // then round as the reference's.  Measured on the CPU model of this arithmetic (tools/arith_model.c, variant 12): 0.1 %
    // of the stored values differ from the reference per iteration, by one ulp, against 6-30 % for the round-1 FAST
    // (variant 4), and the long-run error stays below 4.5e-6 at 128^3 x 300 iterations where that one reached 1.1e-5.
    __device__ __forceinline__ static void runF32(float (&p)[19], const float omega)
    {
        using namespace exact;
        const float X_M1 = fadd(fadd(fadd(fadd(p[0], p[3]), p[4]), p[5]), p[6]);
        const float X_P1 = fadd(fadd(fadd(fadd(p[10], p[13]), p[14]), p[15]), p[16]);
        const float X_0 = fadd(fadd(fadd(fadd(fadd(fadd(fadd(fadd(p[9], p[1]), p[2]), p[7]), p[8]), p[11]), p[12]), p[17]), p[18]);
        const float Y_M1 = fadd(fadd(fadd(fadd(p[1], p[3]), p[7]), p[8]), p[14]);
        const float Y_P1 = fadd(fadd(fadd(fadd(p[4], p[11]), p[13]), p[17]), p[18]);
        const float Z_M1 = fadd(fadd(fadd(fadd(p[2], p[5]), p[7]), p[16]), p[18]);
        const float Z_P1 = fadd(fadd(fadd(fadd(p[6], p[8]), p[12]), p[15]), p[17]);
        const float rho = fadd(fadd(X_M1, X_P1), X_0);
        float       u0, u1, u2;
        div3(fadd(X_P1, -X_M1), fadd(Y_P1, -Y_M1), fadd(Z_P1, -Z_M1), rho, u0, u1, u2);
        const float nus = -fmul(1.5f, fadd(fadd(fmul(u0, u0), fmul(u1, u1)), fmul(u2, u2)));
        const float cu[9] = {u0, u1, u2, fadd(u0, u1), fadd(u0, -u1), fadd(u0, u2), fadd(u0, -u2), fadd(u1, u2), fadd(u1, -u2)};
        const float om1 = fadd(1.f, -omega);
        // rho * w as hi + lo for w = 1/18, 1/36, 1/3 (w = wh + wl, both floats)
        constexpr float wh18 = (float)(1. / 18.), wl18 = (float)(1. / 18. - (double)wh18);
        constexpr float wh36 = (float)(1. / 36.), wl36 = (float)(1. / 36. - (double)wh36);
        constexpr float wh3 = (float)(1. / 3.), wl3 = (float)(1. / 3. - (double)wh3);
        const float     rh18 = fmul(rho, wh18), rl18 = __fmaf_rn(rho, wl18, __fmaf_rn(rho, wh18, -rh18));
        const float     rh36 = fmul(rho, wh36), rl36 = __fmaf_rn(rho, wl36, __fmaf_rn(rho, wh36, -rh36));
        const float     rh3 = fmul(rho, wh3), rl3 = __fmaf_rn(rho, wl3, __fmaf_rn(rho, wh3, -rh3));
        const float     rx18 = fmul(rh18, 6.f), rx36 = fmul(rh36, 6.f);
#pragma unroll
        for (int g = 0; g < 9; ++g) {
            const float rh = g < 3 ? rh18 : rh36, rl = g < 3 ? rl18 : rl36, rx = g < 3 ? rx18 : rx36, c = cu[g];
            const float s = __fmaf_rn(fmul(4.5f, c), c, __fmaf_rn(-3.f, c, nus));  // -3cu + 4.5cu^2 - usqr
            const float eq = fadd(rh, __fmaf_rn(rh, s, rl));
            const float eqopp = __fmaf_rn(rx, c, eq);
            p[g] = __fmaf_rn(om1, p[g], fmul(omega, eq));
            p[g + 10] = __fmaf_rn(om1, p[g + 10], fmul(omega, eqopp));
        }
        p[9] = __fmaf_rn(om1, p[9], fmul(omega, fadd(rh3, __fmaf_rn(rh3, nus, rl3))));
    }
    __device__ __forceinline__ static void run(T (&p)[19], const T omega)
    {
        if constexpr (sizeof(T) == 4) {
            runF32(p, omega);
            return;
        }
        const T X_M1 = p[0] + p[3] + p[4] + p[5] + p[6];
        const T X_P1 = p[10] + p[13] + p[14] + p[15] + p[16];
        const T X_0 = p[9] + p[1] + p[2] + p[7] + p[8] + p[11] + p[12] + p[17] + p[18];
        const T Y_M1 = p[1] + p[3] + p[7] + p[8] + p[14];
        const T Y_P1 = p[4] + p[11] + p[13] + p[17] + p[18];
        const T Z_M1 = p[2] + p[5] + p[7] + p[16] + p[18];
        const T Z_P1 = p[6] + p[8] + p[12] + p[15] + p[17];
        const T rho = X_M1 + X_P1 + X_0;
        const T inv = rcp(rho);
        const T u0 = (X_P1 - X_M1) * inv;
        const T u1 = (Y_P1 - Y_M1) * inv;
        const T u2 = (Z_P1 - Z_M1) * inv;
        const T base = fma_(T(-1.5), fma_(u0, u0, fma_(u1, u1, u2 * u2)), T(1));  // 1 - usqr
        const T om1 = T(1) - omega;
        const T ro = rho * omega;
        const T rw18 = ro * T(1. / 18.);
        const T rw36 = ro * T(1. / 36.);
        const T cu[9] = {u0, u1, u2, u0 + u1, u0 - u1, u0 + u2, u0 - u2, u1 + u2, u1 - u2};
#pragma unroll
        for (int g = 0; g < 9; ++g) {
            const T rw = g < 3 ? rw18 : rw36;
            const T t = fma_(T(4.5) * cu[g], cu[g], base);
            const T a3 = T(3) * cu[g];
            p[g] = fma_(om1, p[g], rw * (t - a3));
            p[g + 10] = fma_(om1, p[g + 10], rw * (t + a3));
        }
        p[9] = fma_(om1, p[9], ro * T(1. / 3.) * base);
    }
};

// D3Q27 generic BGK exactly as apps/lbmMultiRes/collide.h:311-334 + util.h:47-62 evaluate it (single type T).
template <typename T, int FMAD>
struct CollideD3Q27Ref
{
    static constexpr int Q = 27;
    using Compute = T;
    __device__ __forceinline__ static void run(T (&f)[27], const T omega)
    {
        using L = Lattice<27>;
        T rho = 0;
#pragma unroll
        for (int q = 0; q < 27; ++q)
            rho += f[q];
        T vel[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < 27; ++q) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                vel[d] += f[q] * L::c(q, d);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d)
            vel[d] /= rho;
        const T usqr = (3.0 / 2.0) * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            T cu = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d)
                cu += L::c(q, d) * vel[d];
            cu *= 3.0;
            const T feq = rho * L::w(q) * (1. + cu + 0.5 * cu * cu - usqr);
            f[q] = (1 - omega) * f[q] + omega * feq;
        }
    }
};

template <typename T, int FMAD>
struct CollideD3Q27Fast
{
    static constexpr int Q = 27;
    using Compute = T;
    // fp32: as CollideD3Q19Fast::runF32 — float arithmetic with the reference's rounding points.  With T = float the reference
    // (collide.h:311-334, util.h:47-62) computes rho, vel, usqr and cu in float (repeated here operation for operation; a
    // product by a lattice velocity component is exact), feq in double (double weights and literals) rounded to float — here
    // rho*w carried as two floats — and the relaxation in float without contraction.  CPU model (tools/arith_model.c,
    // variant 13): 3.3e-6 after 100 iterations at 64^3 where the FMA-everywhere version reached 1.4e-5.
    __device__ __forceinline__ static void runF32(float (&f)[27], const float omega)
    {
        using namespace exact;
        using L = Lattice<27>;
        float rho = 0.f, m[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 27; ++q)
            rho = fadd(rho, f[q]);
#pragma unroll
        for (int q = 0; q < 27; ++q) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (L::c(q, d) == 1)
                    m[d] = fadd(m[d], f[q]);
                else if (L::c(q, d) == -1)
                    m[d] = fadd(m[d], -f[q]);
            }
        }
        float u[3];
        div3(m[0], m[1], m[2], rho, u[0], u[1], u[2]);
        const float nus = -fmul(1.5f, fadd(fadd(fmul(u[0], u[0]), fmul(u[1], u[1])), fmul(u[2], u[2])));
        const float om1 = fadd(1.f, -omega);
        // rho * w as hi + lo for the four weight classes
        float rh[4], rl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double wd = k == 0 ? 8.0 / 27.0 : (k == 1 ? 2.0 / 27.0 : (k == 2 ? 1.0 / 54.0 : 1.0 / 216.0));
            const float  wh = (float)wd, wl = (float)(wd - (double)wh);
            rh[k] = fmul(rho, wh);
            rl[k] = __fmaf_rn(rho, wl, __fmaf_rn(rho, wh, -rh[k]));
        }
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            float cu = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (L::c(q, d) == 1)
                    cu = fadd(cu, u[d]);
                else if (L::c(q, d) == -1)
                    cu = fadd(cu, -u[d]);
            }
            cu = fmul(cu, 3.f);
            const int   k = (L::c(q, 0) != 0) + (L::c(q, 1) != 0) + (L::c(q, 2) != 0);
            const float s = __fmaf_rn(fmul(0.5f, cu), cu, fadd(cu, nus));  // cu + cu^2/2 - usqr
            const float feq = fadd(rh[k], __fmaf_rn(rh[k], s, rl[k]));
            f[q] = fadd(fmul(om1, f[q]), fmul(omega, feq));
        }
    }
    __device__ __forceinline__ static void run(T (&f)[27], const T omega)
    {
        using L = Lattice<27>;
        using F = CollideD3Q19Fast<T, FMAD>;
        if constexpr (sizeof(T) == 4) {
            runF32(f, omega);
            return;
        }
        T rho = 0, m[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            rho += f[q];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (L::c(q, d) == 1)
                    m[d] += f[q];
                else if (L::c(q, d) == -1)
                    m[d] -= f[q];
            }
        }
        const T inv = F::rcp(rho);
        const T u[3] = {m[0] * inv, m[1] * inv, m[2] * inv};
        const T base = F::fma_(T(-1.5), F::fma_(u[0], u[0], F::fma_(u[1], u[1], u[2] * u[2])), T(1));
        const T om1 = T(1) - omega;
        const T ro = rho * omega;
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            T cu = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (L::c(q, d) == 1)
                    cu += u[d];
                else if (L::c(q, d) == -1)
                    cu -= u[d];
            }
            cu *= T(3);
            const T t = F::fma_(T(0.5) * cu, cu, cu) + base;  // 1 + cu + cu^2/2 - usqr
            f[q] = F::fma_(om1, f[q], ro * T(L::w(q)) * t);
        }
    }
};

// =============================================================== streaming
// Population q of cell x is pulled from cell x - c_q (LbmTools.h:78-96 / stream.h:28-43).
//
// Two phases so that EVERY global load of a thread is in flight before the first one is consumed: loadOne issues the
// aligned 16-byte row load (and, on the two edge lanes of populations with c_x != 0, the one scalar the warp shuffle
// cannot supply); shiftOne then realises the x shift in registers.
template <class L, int q, typename T, int VEC, bool COH = false>
__device__ __forceinline__ void loadOne(const T* __restrict__ cell0, const DenseArgs& a, const int x0, const int y, const int zm,
                                        const int tx, const int lpr, const bool rowOk, T (&v)[VEC], T& edge)
{
    constexpr int cx = L::c(q, 0), cy = L::c(q, 1), cz = L::c(q, 2);
    const int     ys = y - cy, zs = zm - cz;
    // rows outside the allocation are never dereferenced (an enclosed geometry has no bulk cell there)
    const bool ok = rowOk && (cy == 0 || (unsigned)ys < (unsigned)a.ny) && (cz == 0 || (unsigned)zs < (unsigned)a.nzm);
    const T*   p = cell0 + (q * a.pitch_q - cz * a.pitch_z - (int64_t)cy * a.pitch_y);
    ldPredSel<COH>(p, ok, v);
    edge = T(0);
    if constexpr (cx == 1)
        edge = ldPredSel1<COH>(p - 1, tx == 0 && ok && x0 > 0);
    else if constexpr (cx == -1)
        edge = ldPredSel1<COH>(p + VEC, tx == lpr - 1 && ok && x0 + VEC < a.pitch_y);
}

template <class L, int q, typename T, int VEC>
__device__ __forceinline__ void shiftOne(const int tx, const int lpr, T (&v)[VEC], const T edge)
{
    constexpr int cx = L::c(q, 0);
    if constexpr (cx == 1) {
        T e = __shfl_up_sync(0xffffffffu, v[VEC - 1], 1);
        if (tx == 0)
            e = edge;
#pragma unroll
        for (int i = VEC - 1; i > 0; --i)
            v[i] = v[i - 1];
        v[0] = e;
    } else if constexpr (cx == -1) {
        T e = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (tx == lpr - 1)
            e = edge;
#pragma unroll
        for (int i = 0; i < VEC - 1; ++i)
            v[i] = v[i + 1];
        v[VEC - 1] = e;
    }
}

template <class L, typename T, int VEC, bool COH = false, int... Qs>
__device__ __forceinline__ void loadAll(std::integer_sequence<int, Qs...>, const T* __restrict__ cell0, const DenseArgs& a,
                                        const int x0, const int y, const int zm, const int tx, const int lpr, const bool rowOk,
                                        T (&f)[L::Q][VEC], T (&edge)[L::Q])
{
    (loadOne<L, Qs, T, VEC, COH>(cell0, a, x0, y, zm, tx, lpr, rowOk, f[Qs], edge[Qs]), ...);
}
template <class L, typename T, int VEC, int... Qs>
__device__ __forceinline__ void shiftAll(std::integer_sequence<int, Qs...>, const int tx, const int lpr, T (&f)[L::Q][VEC],
                                         const T (&edge)[L::Q])
{
    (shiftOne<L, Qs, T, VEC>(tx, lpr, f[Qs], edge[Qs]), ...);
}

// ---------------------------------------------------------------- wall fix-up (half-way bounce-back, moving walls)
// Bit q of the cell's mask set <=> the cell at x - c_q is not bulk; then
//   in[q] = f_opp(q)(x) + f_opp(q)(x - c_q)      (LbmTools.h:78-96; stream.h:36-41)
// Both operands live outside the data the streaming step fetched, so they cost a round trip to memory.  All loads of
// one cell are issued together (a cell between two walls aside: q and opp(q) share a slot, the second is fetched late).
template <class L>
struct Pairs
{
    // pair p = (lo, hi = opp(lo)), lo < hi, REST excluded: D3Q19 -> 9 pairs (g, g+10), D3Q27 -> 13 pairs
    static constexpr int N = (L::Q - 1) / 2;
    __host__ __device__ static constexpr int lo(int p)
    {
        int n = 0;
        for (int q = 0; q < L::Q; ++q) {
            if (q != L::REST && q < L::opp(q)) {
                if (n == p)
                    return q;
                ++n;
            }
        }
        return -1;
    }
};

template <class L, typename T, int VEC, int P, bool COH = false>
__device__ __forceinline__ void fixLoad(const T* __restrict__ cell, const DenseArgs& a, const uint32_t m, const int i,
                                        T (&f)[L::Q][VEC], T& tb)
{
    constexpr int q = Pairs<L>::lo(P), o = L::opp(q);
    const bool    bq = (m >> q) & 1u, bo = (m >> o) & 1u;
    // bq: f_o(x) + f_o(x - c_q);   bo (only): f_q(x) + f_q(x - c_o) = f_q(x) + f_q(x + c_q)
    // The first operand lands in the slot it replaces (what was pulled from the wall cell is never used); the second in tb.
    // Predicated volatile PTX loads without a consumer here: with plain loads and `if (bq) f[q][i] = first` ptxas put a
    // select right behind every pair of loads and the pairs of a cell went to memory one after the other (ncu r01m: five
    // serialised round trips per x-wall cell, 1.8x the time of the streaming loads).
    const int64_t dn = L::c(q, 2) * a.pitch_z + (int64_t)L::c(q, 1) * a.pitch_y + L::c(q, 0);
    const T*      src = cell + (bq ? o : q) * a.pitch_q;
    f[q][i] = ldPredKeepSel1<COH>(src, bq, f[q][i]);
    f[o][i] = ldPredKeepSel1<COH>(src, bo && !bq, f[o][i]);
    tb = ldPredSel1<COH>(bq ? src - dn : src + dn, bq || bo);
}
template <class L, typename T, int VEC, int P, bool COH = false>
__device__ __forceinline__ void fixUse(const T* __restrict__ cell, const DenseArgs& a, const uint32_t m, const int i, const T tb,
                                       T (&f)[L::Q][VEC])
{
    constexpr int q = Pairs<L>::lo(P), o = L::opp(q);
    const bool    bq = (m >> q) & 1u, bo = (m >> o) & 1u;
    if (bq)
        f[q][i] = f[q][i] + tb;
    else if (bo)
        f[o][i] = f[o][i] + tb;
    if (bq && bo) {  // walls on both sides of the cell along c_q: the second one is fetched late
        const int64_t dn = L::c(q, 2) * a.pitch_z + (int64_t)L::c(q, 1) * a.pitch_y + L::c(q, 0);
        const T*      src = cell + q * a.pitch_q;
        f[o][i] = ldPredSel1<COH>(src, true) + ldPredSel1<COH>(src + dn, true);
    }
}

template <class L, typename T, int VEC, int P0, int P1, bool COH, int... Ps>
__device__ __forceinline__ void fixChunk(std::integer_sequence<int, Ps...>, const T* __restrict__ cell, const DenseArgs& a,
                                         const uint32_t m, const int i, T (&f)[L::Q][VEC])
{
    constexpr int N = P1 - P0;
    T             tb[N];
    (fixLoad<L, T, VEC, P0 + Ps, COH>(cell, a, m, i, f, tb[Ps]), ...);      // pass 1: every load of the chunk
    (fixUse<L, T, VEC, P0 + Ps, COH>(cell, a, m, i, tb[Ps], f), ...);       // pass 2: use them
}

// CH = pairs per batch: CH temporaries live next to the Q*VEC values
template <class L, typename T, int VEC, int CH, bool COH = false>
__device__ __forceinline__ void fixCell(const T* __restrict__ cell, const DenseArgs& a, const uint32_t m, const int i,
                                        T (&f)[L::Q][VEC])
{
    constexpr int NP = Pairs<L>::N;
    if constexpr (NP > 0 * CH)
        fixChunk<L, T, VEC, 0, (NP < CH ? NP : CH), COH>(std::make_integer_sequence<int, (NP < CH ? NP : CH)>{}, cell, a, m, i, f);
    if constexpr (NP > 1 * CH)
        fixChunk<L, T, VEC, CH, (NP < 2 * CH ? NP : 2 * CH), COH>(std::make_integer_sequence<int, (NP < 2 * CH ? NP : 2 * CH) - CH>{}, cell,
                                                                  a, m, i, f);
    if constexpr (NP > 2 * CH)
        fixChunk<L, T, VEC, 2 * CH, NP, COH>(std::make_integer_sequence<int, NP - 2 * CH>{}, cell, a, m, i, f);
    static_assert(NP <= 3 * CH, "chunking covers three batches");
}

// Populations whose pull source lies across an x face of the box.  A bulk cell at x = 1 whose only non-bulk neighbours
// are the three (D3Q19: five directions, D3Q27: nine) in the plane x = 0 carries exactly the bits of the populations
// with c_x = +1 (side 0); at x = nx-2 those with c_x = -1 (side 1), the opposites of the first set.  The step kernel
// fetches the operands of that fix-up speculatively with its streaming loads; the flag word decides whether they are used.
template <class L>
struct XSet
{
    __host__ __device__ static constexpr int count()
    {
        int n = 0;
        for (int q = 0; q < L::Q; ++q)
            n += L::c(q, 0) == 1;
        return n;
    }
    static constexpr int N = count();
    // k-th population with c_x = +1
    __host__ __device__ static constexpr int q(int k)
    {
        int n = 0;
        for (int i = 0; i < L::Q; ++i) {
            if (L::c(i, 0) == 1) {
                if (n == k)
                    return i;
                ++n;
            }
        }
        return -1;
    }
    __host__ __device__ static constexpr uint32_t mask(int side)
    {
        uint32_t m = 0;
        for (int i = 0; i < L::Q; ++i)
            if (L::c(i, 0) == (side ? -1 : 1))
                m |= 1u << i;
        return m;
    }
};

template <class L, typename T, int VEC, int... Ks>
__device__ __forceinline__ void xFixIssue(std::integer_sequence<int, Ks...>, T* sSlot, const T* __restrict__ cell, const DenseArgs& a,
                                          const int side)
{
    // side 0: in[q] = f_o(x) + f_o(x - c_q), q with c_x = +1, o = opp(q);  side 1: the roles of q and o swap
    constexpr int N = XSet<L>::N;
    const int     sgn = side ? -1 : 1;
    ((cpAsync1(sSlot + (L::Q + Ks) * kStepThreads, cell + (int64_t)(side ? XSet<L>::q(Ks) : L::opp(XSet<L>::q(Ks))) * a.pitch_q),
      cpAsync1(sSlot + (L::Q + N + Ks) * kStepThreads,
               cell + (int64_t)(side ? XSet<L>::q(Ks) : L::opp(XSet<L>::q(Ks))) * a.pitch_q -
                   sgn * (L::c(XSet<L>::q(Ks), 2) * a.pitch_z + (int64_t)L::c(XSet<L>::q(Ks), 1) * a.pitch_y + 1))),
     ...);
}
template <class L, typename T, int VEC, int... Ks>
__device__ __forceinline__ void xFixUse(std::integer_sequence<int, Ks...>, const T* sSlot, const int side, const int i, T (&f)[L::Q][VEC])
{
    constexpr int N = XSet<L>::N;
    // the sum lands in slot q (side 0) or opp(q) (side 1)
    ((side ? (void)(f[L::opp(XSet<L>::q(Ks))][i] = sSlot[(L::Q + Ks) * kStepThreads] + sSlot[(L::Q + N + Ks) * kStepThreads])
           : (void)(f[XSet<L>::q(Ks)][i] = sSlot[(L::Q + Ks) * kStepThreads] + sSlot[(L::Q + N + Ks) * kStepThreads])),
     ...);
}

// Wall fix-ups, collision and stores of the VEC cells one thread owns (shared by the direct and the TMA kernel).
template <class COL, typename T, int VEC, int CH = (sizeof(T) == 4 ? 9 : 5), bool COH = false>
__device__ __forceinline__ void finishCells(const DenseArgs& a, const T* __restrict__ cell0, T* __restrict__ out0,
                                            const uint32_t (&fl)[VEC], const bool special, T (&f)[COL::Q][VEC],
                                            T* __restrict__ peerDst = nullptr, const int64_t peerPitchQ = 0, const int peerDir = 0,
                                            const T* sKeep = nullptr, const int specCell = -1, const int specAdj = -1, const int specSide = 0)
{
    constexpr int Q = COL::Q;
    using L = Lattice<Q>;
    bool allBulk = true, anyBulk = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        allBulk = allBulk && flagIsBulk(fl[i]);
        anyBulk = anyBulk || flagIsBulk(fl[i]);
    }
    if (special) {
        // Non-bulk cells are never updated (LbmTools.h:304).  A thread that owns bulk AND non-bulk cells still leaves
        // through one 16-byte store per population: the non-bulk cells carry the value the output field already holds,
        // fetched here — together with the wall fix-up operands, one round trip for both — into the slots whose pulled
        // values are never used, and kept through the collision by a select.  Nobody else writes those cells.
        // (Storing the bulk cells of such threads one by one instead was measured SLOWER on B200, profiles/r01r: partial
        // sector writes cost more than these loads.)
        // (A branch per cell, not 4 x Q predicated loads: the instructions a special warp executes are what keeps it
        // resident — measured.)
        const bool mixed = anyBulk && !allBulk;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            if (mixed && !flagIsBulk(fl[i])) {
                if (i == specCell) {  // fetched speculatively next to the streaming loads (the kernel's x-face threads)
#pragma unroll
                    for (int q = 0; q < Q; ++q)
                        f[q][i] = sKeep[q * kStepThreads];
                } else {
#pragma unroll
                    for (int q = 0; q < Q; ++q)
                        f[q][i] = ldPredCoherent1(out0 + q * a.pitch_q + i, true, f[q][i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const uint32_t m = fl[i] & kMaskBits;
            if (m != 0 && flagIsBulk(fl[i])) {
                // the cell next to an x face whose only walls are across that face: operands fetched with the streaming loads
                if (i == specAdj && m == (specSide ? XSet<L>::mask(1) : XSet<L>::mask(0)))
                    xFixUse<L, T, VEC>(std::make_integer_sequence<int, XSet<L>::N>{}, sKeep, specSide, i, f);
                else
                    fixCell<L, T, VEC, CH, COH>(cell0 + i, a, m, i, f);
            }
        }
    }

    const typename COL::Compute omega = (typename COL::Compute)a.omega;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        T p[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            p[q] = f[q][i];
        const bool bulk = flagIsBulk(fl[i]);
        if (!BulkOnly<COL>::value || bulk)
            COL::run(p, omega);
#pragma unroll
        for (int q = 0; q < Q; ++q)
            f[q][i] = bulk ? p[q] : f[q][i];
    }

    if (!anyBulk)
        return;
#pragma unroll
    for (int q = 0; q < Q; ++q)
        stVec<T, VEC>(out0 + q * a.pitch_q, f[q]);
    if (peerDst != nullptr) {
        // fused halo update: the populations that cross this z face also go straight into the neighbour's ghost plane
        // (peer memory, NVLink stores) — what nlbm_dense_halo_push would copy after the kernel
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (L::c(q, 2) != 0 && L::c(q, 2) == peerDir)
                stVec<T, VEC>(peerDst + q * peerPitchQ, f[q]);
        }
    }
}

// All warps of a boundary plane report in (on every exit path); the last one publishes the counter the neighbour waits for.
__device__ __forceinline__ void faceArrive(const DenseArgs& a, const int face, const int lane)
{
    __threadfence_system();  // my peer stores are visible system-wide before my warp counts as done
    __syncwarp();
    if (lane == 0) {
        const uint32_t done = atomicAdd(a.counter + face, 1u);
        if (done == a.warpsPerFace - 1) {
            a.counter[face] = 0;  // ready for the next launch
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t*>(a.peerFlag[face]) = a.signalValue;
            __threadfence_system();
        }
    }
}

// =============================================================== the kernel
// Resident blocks per SM asked of ptxas, from the registers the Q x VEC population values of a thread occupy: the kernel is
// latency-bound on its one round trip to HBM, so it wants as many warps as it can have without spilling those values.
__host__ __device__ constexpr int stepMinBlocks(int valueRegs) { return valueRegs <= 40 ? 3 : (valueRegs <= 80 ? 2 : 1); }

// grid  = (ceil(segments/blockDim.y), ceil(ny/blockDim.z), planes of the view), block = (32, SEGS, ROWS)
// PEER: the fused step + face push (nlbm_dense_step_push); a separate instantiation so that the plain kernel carries none of it
// The work of one thread block on one tile (bx, by, bz of the launch grid described above).  COH: every load of the INPUT
// field is an ordinary coherent load (no ld.global.nc, no cp.async prefetch of values that change) — for the multi-iteration
// kernel below, whose input was written by other SMs earlier in the same launch.
template <class COL, typename T, int VEC, bool PEER, bool COH>
__device__ __forceinline__ void stepBody(const DenseArgs& a, const void* __restrict__ fieldIn, void* __restrict__ fieldOut, const void* keepCache,
                                         const unsigned bx, const unsigned by, const unsigned bz)
{
    constexpr int Q = COL::Q;
    using L = Lattice<Q>;
    // a warp covers RPW = 32 / LPR consecutive rows x (LPR * VEC) consecutive cells: the narrower the x extent, the fewer
    // warps touch the x walls and pay the wall round trip (summary-first mode keeps whole-row warps, LPR = 32)
    const int lane = threadIdx.x;
    const int lpr = 1 << a.lprLog2, tx = lane & (lpr - 1), r = lane >> a.lprLog2;
    const int seg = bx * blockDim.y + threadIdx.y;
    const int yw = (by * blockDim.z + threadIdx.z) * (32 >> a.lprLog2);
    const int y = yw + r;
    const int vz = bz;
    // fused face push: the two boundary planes come first (blockIdx.z 0 -> plane 0, 1 -> plane nz-1, then 1, 2, ...) so
    // that the neighbours have their ghost planes long before they start the next iteration
    const int  zl = PEER ? (vz == 0 ? 0 : (vz == 1 ? a.nzLocal - 1 : vz - 1)) : 0;
    const int  zm = PEER ? a.zm0 + zl : a.zm0 + vz + (vz >= a.fold ? a.skip : 0);
    const int  face = PEER && vz < 2 ? vz : -1;
    const bool pushes = PEER && face >= 0 && a.peer[face] != nullptr;
    const int  xw = seg * (lpr * VEC);
    if (xw >= a.nx || yw >= a.ny) {  // warp-uniform
        if (pushes)
            faceArrive(a, face, lane);
        return;
    }
    const bool    rowOk = y < a.ny;
    const int64_t row = (int64_t)zm * a.ny + y;
    const int     chunk0 = seg * VEC;  // summary-first mode: the warp's VEC chunks lie in one summary word (VEC divides 32)
    const int     x0 = xw + tx * VEC;
    const int64_t cellOff = (int64_t)zm * a.pitch_z + (int64_t)y * a.pitch_y + x0;
    const T*      cell0 = reinterpret_cast<const T*>(fieldIn) + cellOff;

    // x-face threads: the thread that owns the wall cell at x = 0 / x = nx-1 next to bulk cells (specCell = its index), and
    // the bulk cell next to it (specAdj, side)
    // (a thread that owns a single cell is never mixed: nothing to keep)
    const int  specCell = (VEC > 1 && a.prefetchXFaces && rowOk && a.nx > VEC) ? (x0 == 0 ? 0 : (x0 + VEC >= a.nx && x0 < a.nx ? a.nx - 1 - x0 : -1)) : -1;
    const bool rowsInside = y >= 1 && y + 1 < a.ny && zm >= 1 && zm + 1 < a.nzm;  // every neighbouring row exists
    int        specAdj = -1, specSide = 0;
    if (VEC > 1 && a.specXFix && rowOk && rowsInside && a.nx > 2 * VEC) {  // 
        if (x0 == 0)
            specAdj = 1;
        else if (x0 <= a.nx - 2 && a.nx - 2 < x0 + VEC) {
            specAdj = a.nx - 2 - x0;
            specSide = 1;
        }
    }
    const bool xface = specCell >= 0 || specAdj >= 0;

    // every load of the thread goes out before anything is consumed
    uint2    s = make_uint2(0u, 0u);
    uint32_t fl[VEC];
    uint32_t mapByte = 0;
    if (a.flagMode == kFlagWords)
        ldFlags<VEC>(a.flags + cellOff, rowOk && (a.experiment != 1), fl);
    else if (a.flagMode == kFlagCellMap) {
        mapByte = ldPredU8(a.cellMap + row * (a.pitch_y >> 2) + (x0 >> 2), rowOk);
        ldFlags<VEC>(a.flags + cellOff, xface, fl);  // x-face threads are special for sure: their flag words travel now
    } else
        s = ldPredU2(a.summary + row * a.wpr + (chunk0 >> 5));
    T f[Q][VEC], edge[Q];
    loadAll<L, T, VEC, COH>(std::make_integer_sequence<int, Q>{}, cell0, a, x0, y, zm, tx, lpr, rowOk, f, edge);

    // The two threads of a row that touch the x faces of the box usually own a wall cell next to bulk cells.  Such a thread
    // keeps the wall cell's values of the OUTPUT field through its 16-byte stores (finishCells); fetched after the flags had
    // arrived they were a dependent DRAM round trip that held x-wall warps for as long again as the streaming loads (ncu
    // r01m: 19 % of all stall samples).  Fetch them now, speculatively and without a destination register: cp.async into
    // this thread's shared-memory slot (same 32-byte sectors, no extra traffic when the guess is right; an L2 prefetch
    // pulled whole 128-byte lines and cost the streaming part 6 %, profiles/r01s).  Nothing is assumed about the geometry:
    // a wrong guess is ignored and any other mixed thread loads late as before.
    extern __shared__ __align__(16) unsigned char sKeepRaw[];
    T*        sKeep = reinterpret_cast<T*>(sKeepRaw) + (threadIdx.z * blockDim.y + threadIdx.y) * 32 + lane;
    if (specCell >= 0) {
        // from the field's x-face cache when the caller maintains one (y-contiguous: no isolated DRAM row activations),
        // else from the wall cell's own rows
        const T*      w = reinterpret_cast<const T*>(fieldOut) + cellOff + specCell;
        int64_t       stride = a.pitch_q;
        if (keepCache != nullptr) {
            stride = (int64_t)a.nzm * a.ny;
            w = reinterpret_cast<const T*>(keepCache) + (x0 == 0 ? 0 : Q * stride) + (int64_t)zm * a.ny + y;
        }
#pragma unroll
        for (int q = 0; q < Q; ++q)
            cpAsync1(sKeep + q * kStepThreads, w + q * stride);
    }
    // The bulk cell next to an x face needs, for the populations that would be pulled out of the wall, f_opp(x) + f_opp(wall
    // cell) (finishCells).  Fetched after the flag word they were the second round trip that every warp touching an x face
    // paid with one active lane (r01t: 22 % of the stall samples sat on the flag consumer).  Their addresses do not depend
    // on the flags: fetch them now as well; finishCells uses them only if the cell's wall bits are exactly the x-face set.
    if (specAdj >= 0)
        xFixIssue<L, T, VEC>(std::make_integer_sequence<int, XSet<L>::N>{}, sKeep, cell0 + specAdj, a, specSide);

    bool special;
    if (a.flagMode == kFlagWords) {
        bool plain = true, bulk = false;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            plain = plain && fl[i] == kPlainBulk;
            bulk = bulk || flagIsBulk(fl[i]);
        }
        if (!__any_sync(0xffffffffu, bulk)) {  // no bulk cell in this warp's segment: nothing to update
            if (pushes)
                faceArrive(a, face, lane);
            return;
        }
        special = !plain;
    } else if (a.flagMode == kFlagCellMap) {
        const uint32_t cm = (1u << VEC) - 1u;
        const uint32_t bulkBits = (mapByte >> (x0 & 3)) & cm, specBits = (mapByte >> (4 + (x0 & 3))) & cm;
        if (!__any_sync(0xffffffffu, bulkBits != 0)) {
            if (pushes)
                faceArrive(a, face, lane);
            return;
        }
        special = bulkBits != 0 && specBits != 0;
        if (special) {
            if (!xface)
                ldFlags<VEC>(a.flags + cellOff, true, fl);  // late: wall-adjacent rows, obstacle surfaces
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                fl[i] = bulkBits ? kPlainBulk : ((uint32_t)NLBM_UNDEFINED << NLBM_FLAG_CLASS_SHIFT);
        }
    } else {
        const uint32_t cm = (1u << VEC) - 1u;
        const uint32_t wbulk = (s.y >> (chunk0 & 31)) & cm;
        if (wbulk == 0) {  // no bulk cell in this warp's segment: nothing to update
            if (pushes)
                faceArrive(a, face, lane);
            return;
        }
        const uint32_t wspec = (s.x >> (chunk0 & 31)) & cm;
        special = (wspec >> ((lane * VEC) >> 5)) & 1u;
        if (special) {
            ldFlags<VEC>(a.flags + cellOff, true, fl);
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                fl[i] = kPlainBulk;
        }
    }

    shiftAll<L, T, VEC>(std::make_integer_sequence<int, Q>{}, tx, lpr, f, edge);
    if (xface)
        cpAsyncWait();  // issued with the streaming loads, which have arrived: free
    if (a.experiment == 3)  // measurement only: flags decide who is updated, but no wall fix-ups and no kept wall values
        special = false;

    if constexpr (PEER) {
        const int fi = pushes ? face : 0;
        T*        peerDst = (pushes && rowOk) ? reinterpret_cast<T*>(a.peer[fi]) + a.peerOff[fi] + (int64_t)y * a.pitch_y + x0 : nullptr;
        finishCells<COL, T, VEC, (sizeof(T) == 4 ? 9 : 5), COH>(a, cell0, reinterpret_cast<T*>(fieldOut) + cellOff, fl, special, f, peerDst, a.peerPitchQ[fi],
                                                                fi == 0 ? -1 : 1, sKeep, specCell, specAdj, specSide);
        if (pushes)
            faceArrive(a, face, lane);
    } else {
        finishCells<COL, T, VEC, (sizeof(T) == 4 ? 9 : 5), COH>(a, cell0, reinterpret_cast<T*>(fieldOut) + cellOff, fl, special, f, nullptr, 0, 0, sKeep,
                                                                specCell, specAdj, specSide);
    }
}

template <class COL, typename T, int VEC, bool PEER>
__global__ void __launch_bounds__(kStepThreads, stepMinBlocks(COL::Q * VEC * (int)sizeof(T) / 4)) k_dense_step(const DenseArgs a)
{
    stepBody<COL, T, VEC, PEER, false>(a, a.in, a.out, a.keepCache, blockIdx.x, blockIdx.y, blockIdx.z);
}

// =============================================================== a chain of dependent launches (small boxes, the default of nlbm_dense_step_n)
// One launch per iteration as in k_dense_step, but iteration t+1 is launched while iteration t still runs (programmatic
// dependent launch: every block of t releases its dependents first thing, so t+1's blocks become resident as t's last blocks
// leave), and the blocks of the first `early` planes — about one chip-load of blocks — do NOT wait for the whole predecessor
// grid: a tile of plane z starts as soon as the tiles of planes z-1, z, z+1 of the previous iteration have finished — those wrote
// every value it reads (RAW) and were the only readers of the cells it overwrites (WAR).  Launch gap and ramp-up of t+1 run in
// the shadow of t's tail, which is most of what a small iteration costs as a kernel of its own.  Blocks of the later planes are
// scheduled when the early ones leave; by then t is over and their griddepcontrol.wait returns at once, so only the early planes
// pay for polling and only planes 0..early (what the early tiles of t+1 depend on) pay for publishing.
// planeDone[z] counts finished tiles of plane z over the whole chain (a plane of iteration t+1 is only touched after the same
// plane of t is complete, so the count is monotone in t).
// Cannot deadlock: t+1 is launched only after EVERY block of t has started (that is what releases it), so whatever a block of
// t+1 waits for — counters or the end of t — is resident or finished; by induction the same holds for t and t-1.  The grid of
// the last iteration completes only after all earlier ones have, so the stream sees the chain as one piece of work.
// Loads of the input field are ordinary coherent loads: behind griddepcontrol.wait, or behind one acquire fence per early block
// (which also drops whatever the SM's L1 kept of the field two iterations ago; a line fetched after that fence belongs to a
// complete plane and stays valid until the tiles of t+2 rewrite it, which wait for this block).
template <class COL, typename T, int VEC>
__global__ void __launch_bounds__(kStepThreads, stepMinBlocks(COL::Q * VEC * (int)sizeof(T) / 4)) k_dense_chain(const DenseArgs a, const ChainArgs c)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool first = threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0;
    if (c.target != 0) {
        if (blockIdx.z < c.early) {
            if (first) {
                const unsigned  z = blockIdx.z, zl = z > 0 ? z - 1 : 0, zh = z + 1 < gridDim.z ? z + 1 : z;
                const unsigned *p0 = c.planeDone + zl, *p1 = c.planeDone + z, *p2 = c.planeDone + zh;
                unsigned        v0, v1, v2;
                for (;;) {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(p0) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(p1) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(p2) : "memory");
                    if (v0 >= c.target && v1 >= c.target && v2 >= c.target)
                        break;
                    __nanosleep(40);
                }
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            __syncthreads();
        } else {
            asm volatile("griddepcontrol.wait;" ::: "memory");
        }
    }
    stepBody<COL, T, VEC, false, true>(a, a.in, a.out, a.keepCache, blockIdx.x, blockIdx.y, blockIdx.z);
    if (blockIdx.z <= c.early) {  // what the early tiles of the next iteration wait for
        __syncthreads();          // every store of the block is ordered before the arrival below (bar.sync + the fence's cumulativity)
        if (first) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(c.planeDone + blockIdx.z) : "memory");
        }
    }
}

// Thread-block and grid shape of the direct kernel for one view (shared by the single-iteration and the multi-iteration launch)
template <class COL, typename T, int VEC>
inline void stepGeometry(DenseArgs& a, int nzView, int rowsLog2, int rpwSel, dim3& block, dim3& grid)
{
    // warp tile: RPW rows x (32 / RPW * VEC) cells.  Measured on B200 (profiles/r01k): whole-row warps are best when rows
    // are long (the SoA planes like long contiguous runs); on short rows every whole-row warp touches an x wall and pays
    // the wall round trip, so the tile narrows until a row has at least four pieces — never below 128 bytes per piece.
    // Whole-row warps when the row summary gates the flag loads.
    int lprLog2 = 5;
    if (a.flagMode != kFlagSummaryFirst) {
        int want = 32;
        if (rpwSel > 0) {
            want = 32 >> (rpwSel - 1);
        } else {
            const int minLanes = 128 / (VEC * (int)sizeof(T)) < 2 ? 2 : 128 / (VEC * (int)sizeof(T));
            while (want > minLanes && want * VEC * 4 > a.nx)
                want >>= 1;
        }
        if (want > 32)
            want = 32;
        if (want < 2)
            want = 2;
        lprLog2 = 0;
        while ((1 << lprLog2) < want)
            ++lprLog2;
    }
    a.lprLog2 = lprLog2;
    const int rpw = 32 >> lprLog2;
    const int segs = (a.nx + (VEC << lprLog2) - 1) / (VEC << lprLog2);
    // Block shape: sx warps side by side in x, rows of them stacked in y.  D3Q19 fp32 (16-byte accesses): up to 4 x 2 — a block
    // then reads, per population, two contiguous 2 KB stretches instead of eight of 512 bytes, which the DRAM pages like better
    // (measured on B200, profiles/r02f_rows_sweep.log: 512^3 42.28 -> 43.24 GLUPS, 1024x1024x128 42.74 -> 43.16, 256^3 39.2 ->
    // 41.1, 128^3 30.0 -> 33.7; while the x-wall fix-ups still cost a dependent round trip the stacked shape was the better
    // one, round 1).  Never wider than the row has pieces (64^3: two).  The other lattices / precisions keep the 1 x 8 stack
    // (D3Q27 fp64: 13.93 vs 13.12 GLUPS).
    int warps = kStepThreads / 32;
    int sx = 1;
    if (COL::Q == 19 && sizeof(T) == 4 && VEC == 4) {
        while (sx < 4 && 2 * sx <= segs)
            sx *= 2;
    }
    int rows = warps / sx;
    if (rowsLog2 > 0) {
        rows = 1 << (rowsLog2 - 1);
        if (rows > warps)
            rows = warps;
        sx = warps / rows;
    }
    block = dim3(32, sx, rows);
    grid = dim3((segs + sx - 1) / sx, (a.ny + rows * rpw - 1) / (rows * rpw), nzView > 0 ? nzView : 1);
    a.warpsPerFace = grid.x * grid.y * (unsigned)warps;  // every warp launched for a plane reports in
}
// one slot of Q kept values + 2 x |XSet| fix-up operands per thread for the speculative fetches (only x-face threads use theirs)
template <class COL, typename T, int VEC>
constexpr size_t stepKeepBytes()
{
    return VEC > 1 ? (size_t)(COL::Q + 2 * XSet<Lattice<COL::Q>>::N) * kStepThreads * sizeof(T) : 0;
}

template <class COL, typename T, int VEC, bool PEER = false>
inline cudaError_t launchStepVec(DenseArgs a, int nzView, int rowsLog2, int rpwSel, cudaStream_t st)
{
    dim3 block, grid;
    stepGeometry<COL, T, VEC>(a, nzView, rowsLog2, rpwSel, block, grid);
    if (nzView <= 0)
        return cudaSuccess;
    if (grid.y > 65535)
        return cudaErrorInvalidConfiguration;
    constexpr size_t keepBytes = stepKeepBytes<COL, T, VEC>();
    if constexpr (keepBytes > 48 * 1024) {
        static bool raised[64] = {};  // per instantiation and device; racing host threads set the same value
        int         dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
            return cudaErrorInvalidDevice;
        if (!raised[dev]) {
            cudaError_t e = cudaFuncSetAttribute(k_dense_step<COL, T, VEC, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)keepBytes);
            if (e != cudaSuccess)
                return e;
            raised[dev] = true;
        }
    }
    k_dense_step<COL, T, VEC, PEER><<<grid, block, keepBytes, st>>>(a);
    return cudaGetLastError();
}

// Plane counters of a launch chain: kChainPlanes words of a per-device pool (1 MB, allocated on first use, kept for the life of the
// process).  Chains issued directly on a stream use that stream's own slice — chains on one stream are ordered by the stream (the
// first launch of a chain is an ordinary one), so nothing else can touch it; up to 16 streams per device.  A chain that is being
// captured into a CUDA graph gets a slice of its own for good (the graph may be replayed on any stream at any time); 48 of them.
// No slice left, or the pool cannot be allocated now (first use inside a capture): nullptr — the caller issues plain launches.
constexpr int kChainPlanes = kChainPlanesApi;
inline unsigned* chainPlaneCounters(int dev, cudaStream_t st)
{
    constexpr int kStreams = 16, kGraphs = 48;
    struct Pool
    {
        unsigned*    base = nullptr;
        cudaStream_t owner[kStreams] = {};
        bool         used[kStreams] = {};
        int          graphs = 0;
    };
    static std::mutex mu;
    static Pool       pool[64];
    if (dev < 0 || dev >= 64)
        return nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess)
        return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    Pool&                       p = pool[dev];
    if (p.base == nullptr) {
        if (cap != cudaStreamCaptureStatusNone)
            return nullptr;  // cudaMalloc would invalidate the capture
        if (cudaMalloc(reinterpret_cast<void**>(&p.base), (size_t)(kStreams + kGraphs) * kChainPlanes * sizeof(unsigned)) != cudaSuccess) {
            p.base = nullptr;
            (void)cudaGetLastError();
            return nullptr;
        }
    }
    if (cap != cudaStreamCaptureStatusNone)
        return p.graphs < kGraphs ? p.base + (size_t)(kStreams + p.graphs++) * kChainPlanes : nullptr;
    for (int i = 0; i < kStreams; ++i) {
        if (p.used[i] && p.owner[i] == st)
            return p.base + (size_t)i * kChainPlanes;
    }
    for (int i = 0; i < kStreams; ++i) {
        if (!p.used[i]) {
            p.used[i] = true;
            p.owner[i] = st;
            return p.base + (size_t)i * kChainPlanes;
        }
    }
    return nullptr;
}

// `iterations` iterations as a chain of dependent launches (k_dense_chain); a.in / m.fieldB are the two fields
template <class COL, typename T, int VEC>
inline cudaError_t launchChainVec(DenseArgs a, const MultiArgs& m, int nzView, int rowsLog2, int rpwSel, cudaStream_t st)
{
    dim3 block, grid;
    stepGeometry<COL, T, VEC>(a, nzView, rowsLog2, rpwSel, block, grid);
    if (nzView <= 0 || m.iterations <= 0)
        return cudaSuccess;
    if (grid.y > 65535 || grid.z > (unsigned)kChainPlanes)
        return cudaErrorInvalidConfiguration;
    constexpr size_t keepBytes = stepKeepBytes<COL, T, VEC>();
    int              dev = 0;
    cudaError_t      e = cudaGetDevice(&dev);
    if (e == cudaSuccess && keepBytes > 48 * 1024)
        e = cudaFuncSetAttribute(k_dense_chain<COL, T, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)keepBytes);
    if (e != cudaSuccess)
        return e;
    ChainArgs c;
    c.planeDone = chainPlaneCounters(dev, st);
    if (c.planeDone == nullptr)
        return cudaErrorNotReady;  // no counters to be had: launchMulti issues plain launches instead
    e = cudaMemsetAsync(c.planeDone, 0, grid.z * sizeof(unsigned), st);
    if (e != cudaSuccess)
        return e;
    const void *fieldA = a.in, *fieldB = m.fieldB, *keepB = a.keepCache, *keepA = m.keepCacheA;
    const unsigned tilesPerPlane = grid.x * grid.y;
    // planes whose tiles start on the counters: one chip-load of blocks (what can be resident next to the predecessor's tail)
    static int slots[64] = {};  // per instantiation and device; racing host threads store the same value
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    if (slots[dev] == 0) {
        int sms = 0, perSm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_dense_chain<COL, T, VEC>, kStepThreads, keepBytes);
        if (e != cudaSuccess)
            return e;
        slots[dev] = sms * (perSm > 0 ? perSm : 1);
    }
    c.early = m.chainEarly != 0 ? (unsigned)m.chainEarly : ((unsigned)slots[dev] + tilesPerPlane - 1) / tilesPerPlane + 1;
    if (c.early > grid.z)
        c.early = grid.z;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = keepBytes;
    cfg.stream = st;
    cfg.attrs = attr;
    for (int it = 0; it < m.iterations; ++it) {
        const bool even = (it & 1) == 0;
        a.in = even ? fieldA : fieldB;
        a.out = const_cast<void*>(even ? fieldB : fieldA);
        a.keepCache = even ? keepB : keepA;
        c.target = (unsigned)it * tilesPerPlane;
        cfg.numAttrs = it > 0 ? 1 : 0;  // the first launch of a chain is an ordinary one: behind everything the stream holds
        e = cudaLaunchKernelEx(&cfg, k_dense_chain<COL, T, VEC>, a, c);
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

template <class COL, typename T>
inline cudaError_t launchStep(const DenseArgs& a, int nzView, int vec, int rowsLog2, int rpwSel, cudaStream_t st)
{
    constexpr int maxVec = 16 / (int)sizeof(T);
    if (vec <= 0 || vec > maxVec) {
        // default: the widest access whose values fit two resident blocks per SM (D3Q19: 16 bytes, D3Q27: 8 bytes)
        vec = maxVec;
        while (vec > 1 && COL::Q * vec * (int)sizeof(T) / 4 > 80)
            vec >>= 1;
        // boxes of up to ~one chip-load of cells (a 96^3 box is 3 waves) are bound by the latency of ONE thread's chain of
        // loads, collisions and stores, not by bandwidth: half as many cells per thread, three blocks per SM
        // (profiles/r02i_small_sweep.log: 64^3 22.1 -> 23.4 GLUPS, 96^3 25.4 -> 27.3; 128^3 is better off with 16 bytes)
        if (!a.peerMode && vec == 4 && (int64_t)a.nx * a.ny * nzView <= (int64_t)1 << 20)
            vec = 2;
    }
    while (vec > 1 && (a.pitch_y % (32 * vec) != 0))
        vec >>= 1;
    if (a.peerMode) {  // the fused step + face push exists for the default access width of each lattice / precision
        constexpr int dv = (COL::Q * maxVec * (int)sizeof(T) / 4 > 80) ? maxVec / 2 : maxVec;
        if (a.pitch_y % (32 * dv) != 0)
            return cudaErrorInvalidConfiguration;
        return launchStepVec<COL, T, dv, true>(a, nzView, rowsLog2, rpwSel, st);
    }
    if constexpr (maxVec >= 4) {
        if (vec == 4)
            return launchStepVec<COL, T, 4>(a, nzView, rowsLog2, rpwSel, st);
    }
    if (vec >= 2)
        return launchStepVec<COL, T, 2>(a, nzView, rowsLog2, rpwSel, st);
    return launchStepVec<COL, T, 1>(a, nzView, rowsLog2, rpwSel, st);
}

// `iterations` ordinary step launches, the two fields swapping roles
template <class COL, typename T>
inline cudaError_t launchPlainSteps(const DenseArgs& a, const MultiArgs& m, int nzView, int vec, int rowsLog2, int rpwSel, cudaStream_t st)
{
    DenseArgs   b = a;
    const void *fieldA = a.in, *keepB = a.keepCache;
    for (int it = 0; it < m.iterations; ++it) {
        const bool even = (it & 1) == 0;
        b.in = even ? fieldA : m.fieldB;
        b.out = const_cast<void*>(even ? m.fieldB : fieldA);
        b.keepCache = even ? keepB : m.keepCacheA;
        cudaError_t e = launchStep<COL, T>(b, nzView, vec, rowsLog2, rpwSel, st);
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

// nlbm_dense_step_n: the launch chain; a view with more planes than a chain has counters, or a call that finds no counters free,
// runs as plain step launches
template <class COL, typename T>
inline cudaError_t launchMulti(const DenseArgs& a, const MultiArgs& m, int nzView, int vec, int rowsLog2, int rpwSel, cudaStream_t st)
{
    if (nzView > kChainPlanes)
        return launchPlainSteps<COL, T>(a, m, nzView, vec, rowsLog2, rpwSel, st);
    constexpr int maxVec = 16 / (int)sizeof(T);
    int           cv = vec;
    if (cv <= 0 || cv > maxVec) {
        cv = maxVec;
        while (cv > 1 && COL::Q * cv * (int)sizeof(T) / 4 > 80)
            cv >>= 1;
    }
    while (cv > 1 && (a.pitch_y % (32 * cv) != 0))
        cv >>= 1;
    cudaError_t e;
    if (maxVec >= 4 && cv == 4) {
        if constexpr (maxVec >= 4)
            e = launchChainVec<COL, T, 4>(a, m, nzView, rowsLog2, rpwSel, st);
        else
            e = cudaErrorInvalidValue;
    } else if (cv >= 2) {
        e = launchChainVec<COL, T, 2>(a, m, nzView, rowsLog2, rpwSel, st);
    } else {
        e = launchChainVec<COL, T, 1>(a, m, nzView, rowsLog2, rpwSel, st);
    }
    if (e == cudaErrorNotReady)
        return launchPlainSteps<COL, T>(a, m, nzView, vec, rowsLog2, rpwSel, st);
    return e;
}

}  // namespace nlbm
