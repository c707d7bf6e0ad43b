/* TEST INFRASTRUCTURE — CPU oracle, "template" body.  Included once per
 * (STORE, COMPUTE) pair by lbm_oracle.c with
 *     #define STORE   float|double      storage type of the populations
 *     #define COMPUTE float|double      the reference's LbmComputeType
 *     #define SFX(name) name##_f32 ...
 *
 * Every expression below keeps the operand TYPES and the association order of
 * the reference so that C's usual arithmetic conversions reproduce its
 * rounding exactly: the reference writes its constants as `double` literals,
 * hence with COMPUTE=float most collision expressions are evaluated in double
 * and rounded once at the assignment, while the
 * moments (row a4) are pure COMPUTE arithmetic.  Compile with
 * -ffp-contract=off (see oracle/Makefile).
 *
 * Dense layout used by the oracle: pop[q][z][y][x], cls[z][y][x] (int32,
 * 0 bounceBack, 1 movingWall, 2 bulk), mask[z][y][x] (uint32).
 */

/* ---- D3Q19 pull stream: benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h:78-96,99-168
 * For the pair (g, b=g+10), c_b = -c_g:
 *   bit g set (cell at x - c_g = x + c_b is not bulk):  in[g] = f(x,b) + f(x+c_b, b)
 *   else                                              :  in[g] = f(x+c_b, g)
 * and symmetrically for b.  The sum is formed in STORE type.            */
static void SFX(olbm_d3q19_pull)(const STORE* fin, size_t cells, int nx, int ny,
                                 int x, int y, int z, uint32_t bits, STORE in[19])
{
    const size_t o = ((size_t)z * ny + y) * nx + x;
    for (int g = 0; g < 9; ++g) {
        const int    b = g + 10;
        const int*   cg = OLBM_C19[g];
        const size_t og = ((size_t)(z + cg[2]) * ny + (y + cg[1])) * nx + (x + cg[0]); /* x + c_g */
        const size_t ob = ((size_t)(z - cg[2]) * ny + (y - cg[1])) * nx + (x - cg[0]); /* x + c_b */
        if (bits & (1u << g)) {
            in[g] = fin[(size_t)b * cells + o] + fin[(size_t)b * cells + ob];
        } else {
            in[g] = fin[(size_t)g * cells + ob];
        }
        if (bits & (1u << b)) {
            in[b] = fin[(size_t)g * cells + o] + fin[(size_t)g * cells + og];
        } else {
            in[b] = fin[(size_t)b * cells + og];
        }
    }
    in[9] = fin[(size_t)9 * cells + o];
}

/* ---- D3Q19 moments: LbmTools.h:172-195 (pure COMPUTE arithmetic, this association) */
static void SFX(olbm_d3q19_moments)(const STORE p[19], COMPUTE* rho, COMPUTE u[3])
{
#define P(i) ((COMPUTE)p[i])
    const COMPUTE X_M1 = P(0) + P(3) + P(4) + P(5) + P(6);
    const COMPUTE X_P1 = P(10) + P(13) + P(14) + P(15) + P(16);
    const COMPUTE X_0 = P(9) + P(1) + P(2) + P(7) + P(8) + P(11) + P(12) + P(17) + P(18);
    const COMPUTE Y_M1 = P(1) + P(3) + P(7) + P(8) + P(14);
    const COMPUTE Y_P1 = P(4) + P(11) + P(13) + P(17) + P(18);
    const COMPUTE Z_M1 = P(2) + P(5) + P(7) + P(16) + P(18);
    const COMPUTE Z_P1 = P(6) + P(8) + P(12) + P(15) + P(17);
#undef P
    *rho = X_M1 + X_P1 + X_0;
    u[0] = (X_P1 - X_M1) / *rho;
    u[1] = (Y_P1 - Y_M1) / *rho;
    u[2] = (Z_P1 - Z_M1) / *rho;
}

/* ---- D3Q19 BGK: LbmTools.h:199-282.  cu for pair g: 0:u0 1:u1 2:u2 3:u0+u1 4:u0-u1
 * 5:u0+u2 6:u0-u2 7:u1+u2 8:u1-u2 (COMPUTE arithmetic); weights are the
 * double literals 1./18. (g<3) and 1./36.                                         */
static void SFX(olbm_d3q19_collide)(const STORE p[19], COMPUTE rho, const COMPUTE u[3],
                                    COMPUTE omega, STORE out[19])
{
    const COMPUTE usqr = 1.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]); /* LbmTools.h:312-314 */
    COMPUTE       cu[9];
    cu[0] = u[0];
    cu[1] = u[1];
    cu[2] = u[2];
    cu[3] = u[0] + u[1];
    cu[4] = u[0] - u[1];
    cu[5] = u[0] + u[2];
    cu[6] = u[0] - u[2];
    cu[7] = u[1] + u[2];
    cu[8] = u[1] - u[2];
    for (int g = 0; g < 9; ++g) {
        const double  w = g < 3 ? (1. / 18.) : (1. / 36.);
        const COMPUTE eq = rho * w * (1. - 3. * cu[g] + 4.5 * cu[g] * cu[g] - usqr);
        const COMPUTE eqopp = eq + rho * w * 6. * cu[g];
        const COMPUTE o_go = (1. - omega) * (COMPUTE)p[g] + omega * eq;
        const COMPUTE o_bk = (1. - omega) * (COMPUTE)p[g + 10] + omega * eqopp;
        out[g] = (STORE)o_go;
        out[g + 10] = (STORE)o_bk;
    }
    {
        const COMPUTE eq9 = rho * (1. / 3.) * (1. - usqr);
        const COMPUTE o9 = (1. - omega) * (COMPUTE)p[9] + omega * eq9;
        out[9] = (STORE)o9;
    }
}

/* ---- one iteration, D3Q19: LbmTools.h:285-325.  Non-bulk cells are never written. */
void SFX(olbm_d3q19_step)(int nx, int ny, int nz, const STORE* fin, STORE* fout,
                          const int32_t* cls, const uint32_t* mask, double omega_d)
{
    const size_t  cells = (size_t)nx * ny * nz;
    const COMPUTE omega = (COMPUTE)omega_d; /* Config.h:62 getLbmParameters<ComputeFP> */
    /* cells are independent (two fields): threads over z give the same bits, only sooner (long-run parity tests) */
#pragma omp parallel for schedule(static) if (g_olbm_threads > 1) num_threads(g_olbm_threads > 1 ? g_olbm_threads : 1)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t o = ((size_t)z * ny + y) * nx + x;
                if (cls[o] != OLBM_BULK)
                    continue;
                STORE   in[19], out[19];
                COMPUTE rho, u[3];
                SFX(olbm_d3q19_pull)(fin, cells, nx, ny, x, y, z, mask[o], in);
                SFX(olbm_d3q19_moments)(in, &rho, u);
                SFX(olbm_d3q19_collide)(in, rho, u, omega, out);
                for (int q = 0; q < 19; ++q)
                    fout[(size_t)q * cells + o] = out[q];
            }
}

/* ---- D3Q27 (north-star config 5).  The reference never runs D3Q27 on dGrid;
 * this combines apps/lbmMultiRes/{lattice.h:15-77, stream.h:5-49, collide.h:286-354,
 * util.h:47-62} with the two-population fused pull kernel above ("parity
 * unpinned" upstream).  T there is a single type: STORE==COMPUTE. */
void SFX(olbm_d3q27_step)(int nx, int ny, int nz, const STORE* fin, STORE* fout,
                          const int32_t* cls, const uint32_t* mask, double omega_d)
{
    const size_t cells = (size_t)nx * ny * nz;
    const STORE  omega = (STORE)omega_d;
#pragma omp parallel for schedule(static) if (g_olbm_threads > 1) num_threads(g_olbm_threads > 1 ? g_olbm_threads : 1)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t o = ((size_t)z * ny + y) * nx + x;
                if (cls[o] != OLBM_BULK)
                    continue;
                const uint32_t bits = mask[o];
                STORE          ins[27];
                /* stream.h:28-43: neighbour at -c_q bulk -> pull, else f(x,opp)+f(ngh,opp) */
                for (int q = 0; q < 27; ++q) {
                    const int*   c = OLBM_C27[q];
                    const size_t on = ((size_t)(z - c[2]) * ny + (y - c[1])) * nx + (x - c[0]);
                    if (q != 0 && (bits & (1u << q))) {
                        const int oq = OLBM_OPP27[q];
                        ins[q] = fin[(size_t)oq * cells + o] + fin[(size_t)oq * cells + on];
                    } else {
                        ins[q] = fin[(size_t)q * cells + on];
                    }
                }
                /* collide.h:311-334 */
                STORE rho = 0;
                for (int q = 0; q < 27; ++q)
                    rho += ins[q];
                STORE vel[3] = {0, 0, 0}; /* util.h:47-62 */
                for (int q = 0; q < 27; ++q) {
                    const STORE f = ins[q];
                    for (int d = 0; d < 3; ++d)
                        vel[d] += f * OLBM_C27[q][d];
                }
                for (int d = 0; d < 3; ++d)
                    vel[d] /= rho;
                const STORE usqr = (3.0 / 2.0) * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
                for (int q = 0; q < 27; ++q) {
                    STORE cu = 0;
                    for (int d = 0; d < 3; ++d)
                        cu += OLBM_C27[q][d] * vel[d];
                    cu *= 3.0;
                    const STORE feq = rho * OLBM_W27[q] * (1. + cu + 0.5 * cu * cu - usqr);
                    fout[(size_t)q * cells + o] = (1 - omega) * ins[q] + omega * feq;
                }
            }
}

/* ---- cavity (+ optional obstacle mask) initial populations:
 * RunCavityTwoPop.cu:159-206 (D3Q19) / apps/lbmMultiRes/lidDrivenCavity.h:56-76 (D3Q27).
 * cls must already be filled (olbm_classify_*).  ulid = (1,0,0) for D3Q19 scaled by ulb in
 * the expression; for D3Q27 the wall velocity vector is (ulb,0,0).                      */
void SFX(olbm_init_pop)(int q_lat, int nx, int ny, int nz, double ulb, const int32_t* cls, STORE* pop)
{
    const size_t cells = (size_t)nx * ny * nz;
    for (int q = 0; q < q_lat; ++q)
        for (size_t o = 0; o < cells; ++o) {
            STORE val;
            if (q_lat == 19) {
                const double t = OLBM_W19[q];
                val = t;
                if (cls[o] == OLBM_MOVING) {
                    val = -6. * t * ulb * (OLBM_C19[q][0] * 1. + OLBM_C19[q][1] * 0. + OLBM_C19[q][2] * 0.);
                } else if (cls[o] != OLBM_BULK) {
                    val = 0;
                }
            } else {
                const double uw[3] = {ulb, 0., 0.};
                val = OLBM_W27[q];
                if (cls[o] == OLBM_MOVING) {
                    val = 0;
                    for (int d = 0; d < 3; ++d)
                        val += OLBM_C27[q][d] * uw[d];
                    val *= -6. * OLBM_W27[q];
                } else if (cls[o] != OLBM_BULK) {
                    val = 0;
                }
            }
            pop[(size_t)q * cells + o] = val;
        }
}

/* ---- rho/u output: LbmTools.h:384-437 (D3Q19 only) */
void SFX(olbm_d3q19_rho_u)(int nx, int ny, int nz, const STORE* fin, const int32_t* cls,
                           const uint32_t* mask, STORE* rho_out, STORE* u_out)
{
    const size_t cells = (size_t)nx * ny * nz;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t o = ((size_t)z * ny + y) * nx + x;
                COMPUTE      rho = 0, u[3] = {0, 0, 0};
                STORE        in[19];
                if (cls[o] == OLBM_BULK) {
                    SFX(olbm_d3q19_pull)(fin, cells, nx, ny, x, y, z, mask[o], in);
                    SFX(olbm_d3q19_moments)(in, &rho, u);
                } else if (cls[o] == OLBM_MOVING) {
                    rho = 1.0;
                    for (int d = 0; d < 3; ++d)
                        u[d] = (COMPUTE)fin[(size_t)d * cells + o] / (COMPUTE)(6. * 1. / 18.);
                }
                rho_out[o] = (STORE)rho;
                for (int d = 0; d < 3; ++d)
                    u_out[(size_t)d * cells + o] = (STORE)u[d];
            }
}
