// lbm_step.cuh — the fused pull-stream + BGK collide kernel for dense (dGrid) partitions.
//
// Replaces the reference's generic lambda kernel
//   denseSpan::launchLambdaOnSpanCUDA  (libNeonSet/include/Neon/set/LambdaExecutor.h:12-39)
// carrying LbmContainers::iteration   (benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h:285-325)
// = pullStream (:99-168) + macroscopic (:172-195) + collideBgkUnrolled (:199-282), and for D3Q27
// apps/lbmMultiRes/{stream.h:5-49, collide.h:286-354, util.h:47-62}.
//
// Design (B200, HBM-bound: every population element is read once and written once per iteration):
//  * SoA planes with a 512-byte aligned row pitch; one warp owns 32*VEC consecutive cells of one row;
//    every population row is fetched with ONE aligned 16-byte load per thread.  The +-1 x shift of the
//    populations with c_x != 0 is served inside the warp by a shuffle; only the two edge lanes issue a
//    4/8-byte load, which hits a sector the neighbouring warp streams anyway.
//  * a per-row chunk summary (lbm_common.cuh) lets warps skip flag loads, wall fix-ups and predicated
//    stores where all cells are plain bulk; warps without bulk cells exit at once.
//  * wall handling (half-way bounce-back with the wall's stored population, moving lid included) is a
//    per-cell fix-up executed only by cells whose wallNghBitflag is non-zero.
//  * results leave through 16-byte streaming stores; non-bulk cells are never written (LbmTools.h:304).
#pragma once
#include <utility>

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
    __device__ __forceinline__ static void run(T (&p)[19], const T omega)
    {
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
    __device__ __forceinline__ static void run(T (&f)[27], const T omega)
    {
        using L = Lattice<27>;
        using F = CollideD3Q19Fast<T, FMAD>;
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
template <class L, int q, typename T, int VEC>
__device__ __forceinline__ void pullOne(const T* __restrict__ cell0, const DenseArgs& a, const int x0, const int y,
                                        const int zm, const int lane, T (&f)[VEC])
{
    constexpr int cx = L::c(q, 0), cy = L::c(q, 1), cz = L::c(q, 2);
    const int     ys = y - cy, zs = zm - cz;
    // warp-uniform: rows outside the allocation are never dereferenced (an enclosed geometry has no bulk cell there)
    const bool ok = (cy == 0 || (unsigned)ys < (unsigned)a.ny) && (cz == 0 || (unsigned)zs < (unsigned)a.nzm);
    const T*   p = cell0 + (q * a.pitch_q - cz * a.pitch_z - (int64_t)cy * a.pitch_y);
    T          v[VEC];
    if (ok) {
        ldVec<T, VEC>(p, v);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            v[i] = T(0);
    }
    if constexpr (cx == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            f[i] = v[i];
    } else if constexpr (cx == 1) {
        T e = __shfl_up_sync(0xffffffffu, v[VEC - 1], 1);
        if (lane == 0)
            e = (ok && x0 > 0) ? __ldg(p - 1) : T(0);
        f[0] = e;
#pragma unroll
        for (int i = 1; i < VEC; ++i)
            f[i] = v[i - 1];
    } else {
        T e = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (lane == 31)
            e = (ok && x0 + VEC < a.pitch_y) ? __ldg(p + VEC) : T(0);
        f[VEC - 1] = e;
#pragma unroll
        for (int i = 0; i < VEC - 1; ++i)
            f[i] = v[i + 1];
    }
}

// Wall fix-up of population q for one cell: bit q of the mask set <=> the cell at x - c_q is not bulk; then
//   in[q] = f_opp(q)(x) + f_opp(q)(x - c_q)      (LbmTools.h:78-96; stream.h:36-41)
template <class L, int q, typename T>
__device__ __forceinline__ void fixOne(const T* __restrict__ cell, const DenseArgs& a, const uint32_t m, T& fq)
{
    if constexpr (q != L::REST) {
        if (m & (1u << q)) {
            constexpr int o = L::opp(q);
            const T*      po = cell + o * a.pitch_q;
            const int64_t dn = L::c(q, 2) * a.pitch_z + (int64_t)L::c(q, 1) * a.pitch_y + L::c(q, 0);
            fq = __ldg(po) + __ldg(po - dn);
        }
    }
}

template <class L, typename T, int VEC, int... Qs>
__device__ __forceinline__ void pullAll(std::integer_sequence<int, Qs...>, const T* __restrict__ cell0, const DenseArgs& a,
                                        const int x0, const int y, const int zm, const int lane, T (&f)[L::Q][VEC])
{
    (pullOne<L, Qs, T, VEC>(cell0, a, x0, y, zm, lane, f[Qs]), ...);
}
template <class L, typename T, int VEC, int... Qs>
__device__ __forceinline__ void fixAll(std::integer_sequence<int, Qs...>, const T* __restrict__ cell, const DenseArgs& a,
                                       const uint32_t m, const int i, T (&f)[L::Q][VEC])
{
    (fixOne<L, Qs, T>(cell, a, m, f[Qs][i]), ...);
}

// Wall fix-ups, collision and stores of the VEC cells one thread owns (shared by the direct and the TMA kernel).
template <class COL, typename T, int VEC>
__device__ __forceinline__ void finishCells(const DenseArgs& a, const T* __restrict__ cell0, T* __restrict__ out0,
                                            const uint32_t (&fl)[VEC], const bool special, T (&f)[COL::Q][VEC])
{
    constexpr int Q = COL::Q;
    using L = Lattice<Q>;
    if (special) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const uint32_t m = fl[i] & kMaskBits;
            if (m != 0 && flagIsBulk(fl[i]))
                fixAll<L, T, VEC>(std::make_integer_sequence<int, Q>{}, cell0 + i, a, m, i, f);
        }
    }

    const typename COL::Compute omega = (typename COL::Compute)a.omega;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        T p[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            p[q] = f[q][i];
        COL::run(p, omega);
#pragma unroll
        for (int q = 0; q < Q; ++q)
            f[q][i] = p[q];
    }

    bool allBulk = true;
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        allBulk = allBulk && flagIsBulk(fl[i]);
    if (allBulk) {  // the common case, also for wall-adjacent cells: full 16-byte stores
#pragma unroll
        for (int q = 0; q < Q; ++q)
            stVec<T, VEC>(out0 + q * a.pitch_q, f[q]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            if (flagIsBulk(fl[i])) {
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    __stcs(out0 + q * a.pitch_q + i, f[q][i]);
            }
        }
    }
}

// =============================================================== the kernel
// grid  = (ceil(segments/blockDim.y), ceil(ny/blockDim.z), planes of the view), block = (32, SEGS, ROWS)
template <class COL, typename T, int VEC>
__global__ void __launch_bounds__(kStepThreads) k_dense_step(const DenseArgs a)
{
    constexpr int Q = COL::Q;
    using L = Lattice<Q>;
    const int lane = threadIdx.x;
    const int seg = blockIdx.x * blockDim.y + threadIdx.y;
    const int y = blockIdx.y * blockDim.z + threadIdx.z;
    const int vz = blockIdx.z;
    const int zm = a.zm0 + vz + (vz >= a.fold ? a.skip : 0);
    const int xw = seg * (32 * VEC);
    if (xw >= a.nx || y >= a.ny)
        return;  // warp-uniform
    const int64_t  row = (int64_t)zm * a.ny + y;
    const int      chunk0 = seg * VEC;  // the warp's VEC chunks lie in one summary word (VEC divides 32)
    const uint2    s = __ldg(a.summary + row * a.wpr + (chunk0 >> 5));
    const uint32_t cm = (1u << VEC) - 1u;
    const uint32_t wbulk = (s.y >> (chunk0 & 31)) & cm;
    if (wbulk == 0)
        return;  // no bulk cell in this warp's segment: nothing to update
    const uint32_t wspec = (s.x >> (chunk0 & 31)) & cm;
    const int      x0 = xw + lane * VEC;
    const bool     special = (wspec >> ((lane * VEC) >> 5)) & 1u;

    const int64_t cellOff = (int64_t)zm * a.pitch_z + (int64_t)y * a.pitch_y + x0;
    const T*      cell0 = reinterpret_cast<const T*>(a.in) + cellOff;

    uint32_t fl[VEC];
    if (special) {
        using FV = typename Vec<float, VEC>::type;  // same width as VEC uint32
        const FV  t = __ldg(reinterpret_cast<const FV*>(a.flags + cellOff));
        const uint32_t* e = reinterpret_cast<const uint32_t*>(&t);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            fl[i] = e[i];
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            fl[i] = kPlainBulk;
    }

    T f[Q][VEC];
    pullAll<L, T, VEC>(std::make_integer_sequence<int, Q>{}, cell0, a, x0, y, zm, lane, f);

    finishCells<COL, T, VEC>(a, cell0, reinterpret_cast<T*>(a.out) + cellOff, fl, special, f);
}

// =============================================================== host launcher
template <class COL, typename T, int VEC>
inline cudaError_t launchStepVec(const DenseArgs& a, int nzView, int rowsLog2, cudaStream_t st)
{
    const int segs = (a.nx + 32 * VEC - 1) / (32 * VEC);
    int       warps = kStepThreads / 32;
    int       sx = 1;
    while (sx * 2 <= warps && sx < segs)
        sx *= 2;
    int rows = warps / sx;
    if (rowsLog2 > 0) {
        rows = 1 << (rowsLog2 - 1);
        if (rows > warps)
            rows = warps;
        sx = warps / rows;
    }
    dim3 block(32, sx, rows);
    dim3 grid((segs + sx - 1) / sx, (a.ny + rows - 1) / rows, nzView);
    if (nzView <= 0)
        return cudaSuccess;
    k_dense_step<COL, T, VEC><<<grid, block, 0, st>>>(a);
    return cudaGetLastError();
}

template <class COL, typename T>
inline cudaError_t launchStep(const DenseArgs& a, int nzView, int vec, int rowsLog2, cudaStream_t st)
{
    constexpr int maxVec = 16 / (int)sizeof(T);
    if (vec <= 0 || vec > maxVec)
        vec = maxVec;
    while (vec > 1 && (a.pitch_y % (32 * vec) != 0))
        vec >>= 1;
    if constexpr (maxVec >= 4) {
        if (vec == 4)
            return launchStepVec<COL, T, 4>(a, nzView, rowsLog2, st);
    }
    if (vec >= 2)
        return launchStepVec<COL, T, 2>(a, nzView, rowsLog2, st);
    return launchStepVec<COL, T, 1>(a, nzView, rowsLog2, st);
}

}  // namespace nlbm
