// TEST INFRASTRUCTURE — host check of neon_b200/csrc/lbm_collide_exact.cuh: every CONV variant must produce the bits of the
// operand-for-operand transcription of the reference collision (the expressions of oracle/lbm_oracle_impl.h, which are
// pinned to the reference's own dumps) on random cells, including degenerate ones.
//   g++ -O2 -std=c++17 -ffp-contract=off -fopenmp -o /tmp/exact_check tools/exact_check.cpp && /tmp/exact_check [cells]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <omp.h>

#include "../neon_b200/csrc/lbm_collide_exact.cuh"

// LbmTools.h:172-195, 199-282, 312-314 with ComputeFP = Store = float (as oracle/lbm_oracle_impl.h restates them)
static void reference(const float* p, float omega, float* out)
{
    const float X_M1 = p[0] + p[3] + p[4] + p[5] + p[6];
    const float X_P1 = p[10] + p[13] + p[14] + p[15] + p[16];
    const float X_0 = p[9] + p[1] + p[2] + p[7] + p[8] + p[11] + p[12] + p[17] + p[18];
    const float Y_M1 = p[1] + p[3] + p[7] + p[8] + p[14];
    const float Y_P1 = p[4] + p[11] + p[13] + p[17] + p[18];
    const float Z_M1 = p[2] + p[5] + p[7] + p[16] + p[18];
    const float Z_P1 = p[6] + p[8] + p[12] + p[15] + p[17];
    const float rho = X_M1 + X_P1 + X_0;
    const float u0 = (X_P1 - X_M1) / rho, u1 = (Y_P1 - Y_M1) / rho, u2 = (Z_P1 - Z_M1) / rho;
    const float usqr = 1.5 * (u0 * u0 + u1 * u1 + u2 * u2);
    const float cu[9] = {u0, u1, u2, u0 + u1, u0 - u1, u0 + u2, u0 - u2, u1 + u2, u1 - u2};
    for (int g = 0; g < 9; ++g) {
        const double w = g < 3 ? (1. / 18.) : (1. / 36.);
        const float  eq = rho * w * (1. - 3. * cu[g] + 4.5 * cu[g] * cu[g] - usqr);
        const float  eqopp = eq + rho * w * 6. * cu[g];
        out[g] = (1. - omega) * p[g] + omega * eq;
        out[g + 10] = (1. - omega) * p[g + 10] + omega * eqopp;
    }
    const float eq9 = rho * (1. / 3.) * (1. - usqr);
    out[9] = (1. - omega) * p[9] + omega * eq9;
}

template <int CONV>
static void candidate(const float* p, float omega, float* out, long& slow)
{
    float f[19];
    memcpy(f, p, sizeof f);
    if (!nlbm::exact::collideD3Q19<CONV>(f, omega)) {
        ++slow;
        memcpy(f, p, sizeof f);
        nlbm::exact::collideD3Q19<0>(f, omega);
    }
    memcpy(out, f, sizeof f);
}

int main(int argc, char** argv)
{
    const long   n = argc > 1 ? atol(argv[1]) : 20000000;
    const double W[19] = {1. / 18, 1. / 18, 1. / 18, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 3,
                          1. / 18, 1. / 18, 1. / 18, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36};
    long         bad[4] = {0, 0, 0, 0}, slow[4] = {0, 0, 0, 0};
#pragma omp parallel
    {
        std::mt19937_64                        rng(1234 + 77 * omp_get_thread_num());
        std::uniform_real_distribution<double> uni(-1., 1.);
        long                                   lb[4] = {0, 0, 0, 0}, ls[4] = {0, 0, 0, 0};
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            float p[19], ref[19], out[19];
            const int kind = (int)(i % 16);
            const double amp = kind < 8 ? 0.02 : (kind < 12 ? 0.3 : (kind < 14 ? 1e-4 : 0.9));
            for (int q = 0; q < 19; ++q)
                p[q] = (float)(W[q] * (1. + amp * uni(rng)));
            if (kind == 15) {  // degenerate: zeros, exact rest state, a subnormal
                for (int q = 0; q < 19; ++q)
                    p[q] = (float)W[q];
                if (i % 3 == 0)
                    p[(i / 16) % 19] = 0.f;
                if (i % 5 == 0)
                    p[(i / 7) % 19] = 1e-41f;
            }
            const float omega = (float)(0.5 + 1.49 * (0.5 + 0.5 * uni(rng)));
            reference(p, omega, ref);
            candidate<0>(p, omega, out, ls[0]);
            lb[0] += memcmp(ref, out, sizeof ref) != 0;
            candidate<1>(p, omega, out, ls[1]);
            lb[1] += memcmp(ref, out, sizeof ref) != 0;
            candidate<1>(p, omega, out, ls[2]);
            lb[2] += memcmp(ref, out, sizeof ref) != 0;
            candidate<1>(p, omega, out, ls[3]);
            lb[3] += memcmp(ref, out, sizeof ref) != 0;
        }
#pragma omp critical
        for (int v = 0; v < 4; ++v) {
            bad[v] += lb[v];
            slow[v] += ls[v];
        }
    }
    int rc = 0;
    for (int v = 0; v < 4; ++v) {
        printf("CONV=%d: %ld cells, %ld differ from the reference transcription, %ld took the slow path\n", v, n, bad[v], slow[v]);
        rc |= bad[v] != 0;
    }
    return rc;
}
