/* TEST INFRASTRUCTURE / design tool — CPU model of candidate "FAST" arithmetics of the step kernel's collision.
 * Used to choose the arithmetic of Collide*Fast (neon_b200/csrc/lbm_step.cuh) by its error growth against the oracle over
 * long runs, without GPU time.  Same operations as the CUDA code (fmaf/fma, correctly rounded reciprocal), so the curves
 * it produces are the GPU's.  Not part of the product; not used by tests.
 *   gcc -O2 -ffp-contract=off -mfma -fopenmp -fPIC -shared -o /tmp/libarith_model.so tools/arith_model.c -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define BULK 2

static void pull(int Q, const int (*c)[3], const int* opp, int rest, const void* fin_, int isd, size_t cells, int nx, int ny, int x,
                 int y, int z, uint32_t bits, double* in /* values exactly representable in the storage type */)
{
    const size_t o = ((size_t)z * ny + y) * nx + x;
    for (int q = 0; q < Q; ++q) {
        const size_t on = ((size_t)(z - c[q][2]) * ny + (y - c[q][1])) * nx + (x - c[q][0]);
        if (q != rest && (bits >> q & 1u)) {
            const int oq = opp[q];
            if (isd) {
                const double* f = (const double*)fin_;
                in[q] = f[oq * cells + o] + f[oq * cells + on];
            } else {
                const float* f = (const float*)fin_;
                in[q] = (float)(f[oq * cells + o] + f[oq * cells + on]);
            }
        } else {
            in[q] = isd ? ((const double*)fin_)[q * cells + on] : (double)((const float*)fin_)[q * cells + on];
        }
    }
}

#define DEFINE(T, SFX, FMA, ISD)                                                                                                      \
    /* variant 0: the round-1 FAST: sums, reciprocal, FMAs in T */                                                                    \
    static void collide0_##SFX(int Q, const int (*c)[3], const double* w, const T* f, T omega, T* out)                                \
    {                                                                                                                                 \
        T rho = 0, m[3] = {0, 0, 0};                                                                                                  \
        for (int q = 0; q < Q; ++q) {                                                                                                 \
            rho += f[q];                                                                                                              \
            for (int d = 0; d < 3; ++d)                                                                                               \
                m[d] += (T)c[q][d] * f[q];                                                                                            \
        }                                                                                                                             \
        const T inv = (T)1 / rho;                                                                                                     \
        const T u[3] = {m[0] * inv, m[1] * inv, m[2] * inv};                                                                          \
        const T base = FMA((T)-1.5, FMA(u[0], u[0], FMA(u[1], u[1], u[2] * u[2])), (T)1);                                             \
        const T om1 = (T)1 - omega, ro = rho * omega;                                                                                 \
        for (int q = 0; q < Q; ++q) {                                                                                                 \
            T cu = 0;                                                                                                                 \
            for (int d = 0; d < 3; ++d)                                                                                               \
                cu += (T)c[q][d] * u[d];                                                                                              \
            cu *= (T)3;                                                                                                               \
            const T t = FMA((T)0.5 * cu, cu, cu) + base;                                                                              \
            out[q] = FMA(om1, f[q], ro * (T)w[q] * t);                                                                                \
        }                                                                                                                             \
    }                                                                                                                                 \
    /* variant 1: deviations from the rest state: df = f - w, drho = sum df, out = f + omega*(w*(drho + rho*poly) - df) */            \
    static void collide1_##SFX(int Q, const int (*c)[3], const double* w, const T* f, T omega, T* out)                                \
    {                                                                                                                                 \
        T df[27], drho = 0, m[3] = {0, 0, 0};                                                                                         \
        double wres = 0; /* sum over q of (T)w - w: what the rounded weights miss of 1 */                                             \
        for (int q = 0; q < Q; ++q) {                                                                                                 \
            df[q] = f[q] - (T)w[q];                                                                                                   \
            wres += (double)(T)w[q] - w[q];                                                                                           \
        }                                                                                                                             \
        for (int q = 0; q < Q; ++q) {                                                                                                 \
            drho += df[q];                                                                                                            \
            for (int d = 0; d < 3; ++d)                                                                                               \
                m[d] += (T)c[q][d] * df[q];                                                                                           \
        }                                                                                                                             \
        drho += (T)wres;                                                                                                              \
        const T rho = (T)1 + drho;                                                                                                    \
        const T inv = (T)1 / rho;                                                                                                     \
        const T u[3] = {m[0] * inv, m[1] * inv, m[2] * inv};                                                                          \
        const T nusq = (T)-1.5 * FMA(u[0], u[0], FMA(u[1], u[1], u[2] * u[2]));                                                       \
        for (int q = 0; q < Q; ++q) {                                                                                                 \
            T cu = 0;                                                                                                                 \
            for (int d = 0; d < 3; ++d)                                                                                               \
                cu += (T)c[q][d] * u[d];                                                                                              \
            cu *= (T)3;                                                                                                               \
            const T poly = FMA((T)0.5 * cu, cu, cu) + nusq;                                                                           \
            /* (T)w - w folded in: eq - f = w*(drho + rho*poly) - (f - w) , f - w = df + ((T)w - w) */                                \
            const T corr = FMA((T)w[q], FMA(rho, poly, drho), -df[q]) - (T)((double)(T)w[q] - w[q]);                                  \
            out[q] = FMA(omega, corr, f[q]);                                                                                          \
        }                                                                                                                             \
    }

DEFINE(float, f32, fmaf, 0)
DEFINE(double, f64, fma, 1)

/* variant 2 (fp32 storage): everything in double with FMAs, one rounding at the store */
static void collide2_f32(int Q, const int (*c)[3], const double* w, const float* f, float omega_f, float* out)
{
    double       fd[27], o[27];
    const double omega = omega_f;
    for (int q = 0; q < Q; ++q)
        fd[q] = f[q];
    collide0_f64(Q, c, w, fd, omega, o);
    for (int q = 0; q < Q; ++q)
        out[q] = (float)o[q];
}
/* variant 3 (fp32 storage): moments in double, the rest in float as variant 0 */
static void collide3_f32(int Q, const int (*c)[3], const double* w, const float* f, float omega, float* out)
{
    double rho = 0, m[3] = {0, 0, 0};
    for (int q = 0; q < Q; ++q) {
        rho += f[q];
        for (int d = 0; d < 3; ++d)
            m[d] += c[q][d] * (double)f[q];
    }
    const double inv = 1.0 / rho;
    const float  u[3] = {(float)(m[0] * inv), (float)(m[1] * inv), (float)(m[2] * inv)};
    const float  rhof = (float)rho;
    const float  base = fmaf(-1.5f, fmaf(u[0], u[0], fmaf(u[1], u[1], u[2] * u[2])), 1.f);
    const float  om1 = 1.f - omega, ro = rhof * omega;
    for (int q = 0; q < Q; ++q) {
        float cu = 0;
        for (int d = 0; d < 3; ++d)
            cu += (float)c[q][d] * u[d];
        cu *= 3.f;
        const float t = fmaf(0.5f * cu, cu, cu) + base;
        out[q] = fmaf(om1, f[q], ro * (float)w[q] * t);
    }
}


/* ---- D3Q19 fp32 storage, the reference's own moment expressions (LbmTools.h:172-195: float sums in this association,
 * three divisions) so that rho, u, usqr and cu carry the reference's bits; variants differ in the collision only. */
static void moments19(const float* p, float* rho, float* u)
{
    const float X_M1 = p[0] + p[3] + p[4] + p[5] + p[6];
    const float X_P1 = p[10] + p[13] + p[14] + p[15] + p[16];
    const float X_0 = p[9] + p[1] + p[2] + p[7] + p[8] + p[11] + p[12] + p[17] + p[18];
    const float Y_M1 = p[1] + p[3] + p[7] + p[8] + p[14];
    const float Y_P1 = p[4] + p[11] + p[13] + p[17] + p[18];
    const float Z_M1 = p[2] + p[5] + p[7] + p[16] + p[18];
    const float Z_P1 = p[6] + p[8] + p[12] + p[15] + p[17];
    *rho = X_M1 + X_P1 + X_0;
    u[0] = X_P1 - X_M1;
    u[1] = Y_P1 - Y_M1;
    u[2] = Z_P1 - Z_M1;
}
static void collide19(int variant, const float* p, float omega, float* out)
{
    float rho, u[3];
    moments19(p, &rho, u);
    if (variant == 4) { /* the round-1 GPU FAST: reciprocal, FMAs */
        const float inv = 1.f / rho;
        const float u0 = u[0] * inv, u1 = u[1] * inv, u2 = u[2] * inv;
        const float base = fmaf(-1.5f, fmaf(u0, u0, fmaf(u1, u1, u2 * u2)), 1.f);
        const float om1 = 1.f - omega, ro = rho * omega, rw18 = ro * (float)(1. / 18.), rw36 = ro * (float)(1. / 36.);
        const float cu[9] = {u0, u1, u2, u0 + u1, u0 - u1, u0 + u2, u0 - u2, u1 + u2, u1 - u2};
        for (int g = 0; g < 9; ++g) {
            const float rw = g < 3 ? rw18 : rw36;
            const float t = fmaf(4.5f * cu[g], cu[g], base), a3 = 3.f * cu[g];
            out[g] = fmaf(om1, p[g], rw * (t - a3));
            out[g + 10] = fmaf(om1, p[g + 10], rw * (t + a3));
        }
        out[9] = fmaf(om1, p[9], ro * (float)(1. / 3.) * base);
        return;
    }
    /* exact reference moments from here on */
    const float u0 = u[0] / rho, u1 = u[1] / rho, u2 = u[2] / rho;
    const float usqr = 1.5f * (u0 * u0 + u1 * u1 + u2 * u2); /* == (float)(1.5 * (double)(float sum)) */
    const float cu[9] = {u0, u1, u2, u0 + u1, u0 - u1, u0 + u2, u0 - u2, u1 + u2, u1 - u2};
    if (variant == 5) { /* float FMAs */
        const float base = 1.f - usqr;
        const float om1 = 1.f - omega, ro = rho * omega, rw18 = ro * (float)(1. / 18.), rw36 = ro * (float)(1. / 36.);
        for (int g = 0; g < 9; ++g) {
            const float rw = g < 3 ? rw18 : rw36;
            const float t = fmaf(4.5f * cu[g], cu[g], base), a3 = 3.f * cu[g];
            out[g] = fmaf(om1, p[g], rw * (t - a3));
            out[g + 10] = fmaf(om1, p[g + 10], rw * (t + a3));
        }
        out[9] = fmaf(om1, p[9], ro * (float)(1. / 3.) * base);
    } else if (variant == 6) { /* double with FMAs, no intermediate float roundings, narrowing at the store */
        const double r = rho, om = omega, om1 = 1. - om, b = 1. - (double)usqr;
        for (int g = 0; g < 9; ++g) {
            const double w = g < 3 ? (1. / 18.) : (1. / 36.), c = cu[g];
            const double t = fma(4.5 * c, c, b), rw = r * w * om;
            out[g] = (float)fma(om1, (double)p[g], rw * (t - 3. * c));
            out[g + 10] = (float)fma(om1, (double)p[g + 10], rw * (t + 3. * c));
        }
        out[9] = (float)fma(om1, (double)p[9], r * (1. / 3.) * om * b);
    } else if (variant == 7) { /* the reference expressions (sanity: must give zero error) */
        for (int g = 0; g < 9; ++g) {
            const double w = g < 3 ? (1. / 18.) : (1. / 36.);
            const float  eq = rho * w * (1. - 3. * cu[g] + 4.5 * cu[g] * cu[g] - usqr);
            const float  eqopp = eq + rho * w * 6. * cu[g];
            out[g] = (1. - omega) * p[g] + omega * eq;
            out[g + 10] = (1. - omega) * p[g + 10] + omega * eqopp;
        }
        const float eq9 = rho * (1. / 3.) * (1. - usqr);
        out[9] = (1. - omega) * p[9] + omega * eq9;
    } else if (variant == 11) {
        /* float only, but with the reference's rounding points: eq and eqopp rounded to float on their own, omega*eq a float
         * product, the relaxation one FMA (== the reference's double sum rounded to float, double rounding aside) */
        const float om1 = 1.f - omega;
        const float rw18 = rho * (float)(1. / 18.), rw36 = rho * (float)(1. / 36.);
        const float nus = -usqr;
        for (int g = 0; g < 9; ++g) {
            const float rw = g < 3 ? rw18 : rw36, c = cu[g];
            const float s = fmaf(4.5f * c, c, fmaf(-3.f, c, nus)); /* -3c + 4.5c^2 - usqr */
            const float eq = fmaf(rw, s, rw);
            const float eqopp = fmaf(rw * 6.f, c, eq);
            out[g] = fmaf(om1, p[g], omega * eq);
            out[g + 10] = fmaf(om1, p[g + 10], omega * eqopp);
        }
        const float r3 = rho * (float)(1. / 3.);
        out[9] = fmaf(om1, p[9], omega * fmaf(r3, nus, r3));
    } else if (variant == 12) {
        /* as 11 with rho*w carried as two floats (hi + lo): eq is then within ~0.6 ulp of the reference's */
        const float om1 = 1.f - omega;
        const float nus = -usqr;
        float rh[3], rl[3];
        const double wd[3] = {1. / 18., 1. / 36., 1. / 3.};
        for (int k = 0; k < 3; ++k) {
            const float wh = (float)wd[k], wl = (float)(wd[k] - (double)wh);
            rh[k] = rho * wh;
            rl[k] = fmaf(rho, wh, -rh[k]) + rho * wl;
        }
        for (int g = 0; g < 9; ++g) {
            const int   k = g < 3 ? 0 : 1;
            const float c = cu[g];
            const float s = fmaf(4.5f * c, c, fmaf(-3.f, c, nus));
            const float eq = rh[k] + fmaf(rh[k], s, rl[k]);
            const float eqopp = fmaf(rh[k] * 6.f, c, eq);
            out[g] = fmaf(om1, p[g], omega * eq);
            out[g + 10] = fmaf(om1, p[g + 10], omega * eqopp);
        }
        out[9] = fmaf(om1, p[9], omega * (rh[2] + fmaf(rh[2], nus, rl[2])));
    } else if (variant == 8) {
        /* float-float where it matters: eq, eqopp rounded (almost always) as the reference rounds them, the relaxation
         * as one FMA plus the product's error term */
        const float om1 = 1.f - omega; /* exact for omega in [0.5, 2] */
        for (int g = 0; g < 9; ++g) {
            const double wd = g < 3 ? (1. / 18.) : (1. / 36.);
            const float  wh = (float)wd, wl = (float)(wd - (double)wh);
            const float  rh = rho * wh, rl = fmaf(rho, wh, -rh) + rho * wl; /* rho*w = rh + rl */
            const float  c = cu[g];
            /* s = -3c + 4.5c^2 - usqr as hi + lo */
            const float c45 = 4.5f * c;                /* exact? 4.5 = 9/2: 9*c needs 4 more bits: not exact */
            const float c45l = fmaf(4.5f, c, -c45);
            const float q = c45 * c, ql = fmaf(c45, c, -q) + c45l * c; /* 4.5c^2 = q + ql */
            const float a = -3.f * c, al = fmaf(-3.f, c, -a);          /* -3c = a + al */
            /* sum a + q - usqr: magnitudes ~0.1, 0.01, 0.003 */
            const float s1 = a + q, e1 = (a - s1) + q; /* fast two-sum, |a| >= |q| mostly; fall back below */
            const float s2 = s1 - usqr, e2 = (s1 - s2) - usqr;
            const float slo = (e1 + e2) + (al + ql);
            /* eq = (rh + rl) * (1 + s2 + slo) = rh + [rh*s2 + (rl + rh*slo + rl*s2)] */
            const float t = fmaf(rh, s2, fmaf(rh, slo, rl));
            const float eq = rh + t;
            /* eqopp = eq + rho*w*6*cu */
            const float c6 = 6.f * c, c6l = fmaf(6.f, c, -c6);
            const float m = rh * c6, ml = fmaf(rh, c6, -m) + fmaf(rl, c6, rh * c6l);
            const float eqopp = eq + (m + ml);  /* one rounding of eq + m (+ml): approx */
            out[g] = fmaf(om1, p[g], omega * eq);
            out[g + 10] = fmaf(om1, p[g + 10], omega * eqopp);
        }
        {
            const double wd = 1. / 3.;
            const float  wh = (float)wd, wl = (float)(wd - (double)wh);
            const float  rh = rho * wh, rl = fmaf(rho, wh, -rh) + rho * wl;
            const float  eq9 = rh + fmaf(rh, -usqr, rl);
            out[9] = fmaf(om1, p[9], omega * eq9);
        }
    }
}

/* D3Q27 fp32 storage, float only with the reference's rounding points (collide.h:311-334, util.h:47-62 with T = float):
 * rho, vel, usqr, cu are float arithmetic there and are repeated operation for operation; feq is evaluated in double there
 * (double weights and literals) and rounded to float: here rho*w is carried as two floats; the relaxation is float
 * arithmetic without contraction, as the reference's. */
static void collide27_v13(const int (*c)[3], const double* w, const float* f, float omega, float* out)
{
    float rho = 0;
    for (int q = 0; q < 27; ++q)
        rho += f[q];
    float vel[3] = {0, 0, 0};
    for (int q = 0; q < 27; ++q)
        for (int d = 0; d < 3; ++d)
            vel[d] += f[q] * (float)c[q][d];
    for (int d = 0; d < 3; ++d)
        vel[d] /= rho;
    const float usqr = 1.5f * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
    const float om1 = 1 - omega;
    for (int q = 0; q < 27; ++q) {
        float cu = 0;
        for (int d = 0; d < 3; ++d)
            cu += (float)c[q][d] * vel[d];
        cu *= 3.0f;
        const float wh = (float)w[q], wl = (float)(w[q] - (double)wh);
        const float rh = rho * wh, rl = fmaf(rho, wh, -rh) + rho * wl;
        const float s = fmaf(0.5f * cu, cu, cu - usqr);
        const float feq = rh + fmaf(rh, s, rl);
        out[q] = om1 * f[q] + omega * feq;
    }
}

void model_step(int variant, int isd, int Q, const int* c_, const int* opp, const double* w, int nx, int ny, int nz, const void* fin,
                void* fout, const int32_t* cls, const uint32_t* mask, double omega)
{
    const int(*c)[3] = (const int(*)[3])c_;
    const size_t cells = (size_t)nx * ny * nz;
    int          rest = 0;
    for (int q = 0; q < Q; ++q)
        if (!c[q][0] && !c[q][1] && !c[q][2])
            rest = q;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t o = ((size_t)z * ny + y) * nx + x;
                if (cls[o] != BULK)
                    continue;
                double in[27];
                pull(Q, c, opp, rest, fin, isd, cells, nx, ny, x, y, z, mask[o], in);
                if (isd) {
                    double out[27];
                    if (variant == 1)
                        collide1_f64(Q, c, w, in, omega, out);
                    else
                        collide0_f64(Q, c, w, in, omega, out);
                    for (int q = 0; q < Q; ++q)
                        ((double*)fout)[q * cells + o] = out[q];
                } else {
                    float f[27], out[27];
                    for (int q = 0; q < Q; ++q)
                        f[q] = (float)in[q];
                    if (variant == 13 && Q == 27)
                        collide27_v13(c, w, f, (float)omega, out);
                    else if (variant >= 4 && Q == 19)
                        collide19(variant, f, (float)omega, out);
                    else if (variant == 1)
                        collide1_f32(Q, c, w, f, (float)omega, out);
                    else if (variant == 2)
                        collide2_f32(Q, c, w, f, (float)omega, out);
                    else if (variant == 3)
                        collide3_f32(Q, c, w, f, (float)omega, out);
                    else
                        collide0_f32(Q, c, w, f, (float)omega, out);
                    for (int q = 0; q < Q; ++q)
                        ((float*)fout)[q * cells + o] = out[q];
                }
            }
}
