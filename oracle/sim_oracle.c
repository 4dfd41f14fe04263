/* oracle/sim_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C CPU restatement of the hot path of aghaeifar/SpinWalk v1.21.0 `sim`:
 * the per-spin Monte-Carlo time loop  src/sim/kernels.cu:72-233  with its helpers
 * (src/sim/kernels.cuh), parameters::prepare (src/sim/simulation_parameters.cuh:227-245)
 * and the per-scale host driver (src/sim/monte_carlo.cu:236-244,264-337).
 *
 * PARITY PINNING.  This restatement is pinned against the reference itself, compiled in this
 * container from its own untouched source (oracle/Makefile target `ref` -> oracle/_ref/):
 *   flavour SWO_RNG_MT19937 == oracle/_ref/libswref_cpu.so   (g++ build; bit-exact, tests/test_oracle_pin.py)
 *   flavour SWO_RNG_MINSTD  == oracle/_ref/libswref_cuda.so  (nvcc build, host instantiation; bit-exact
 *                               on the committed cases) and the committed fixtures tests/golden/.
 * The reference's own test-suite holds no golden vector for the time loop (tests/test_sim.cu:21-43
 * only checks run()==true); its unit tests of sub2ind / rotations / relax (tests/test_kernel.cpp:17-89)
 * are restated in tests/test_oracle_units.py.
 *
 * Third-party arithmetic the reference leans on and that is NOT under /root/reference:
 *   - libstdc++ 13 <random>: mt19937, generate_canonical<float,24>, Marsaglia polar
 *     normal_distribution<float>  (bits/random.tcc:1809-1844, 3349-3381)        -> restated below
 *   - Thrust/CCCL 2.8 (CUDA 12.9): minstd_rand, normal_distribution_nvcc::sample,
 *     uniform_real_distribution  (thrust/random/detail/*.h, *.inl)              -> restated below
 *   - CUDA host erfcinvf == (float)erfcinv((double)x) (crt/math_functions.hpp:3367) -> swo_erfcinv():
 *     an independent double-precision inverse (Giles' erfinv start + Newton on libm erfc).
 * Each reference line followed is cited as  [file:line].
 */
#include "sim_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SWO_DEG2RAD 0.0174532925199433 /* [simulation_parameters.cuh:24] */
#define SWO_RAD2DEG 57.2957795130823   /* [simulation_parameters.cuh:25] */
#define SWO_GAMMA   267515315.         /* [definitions.h:20] rad/s/T     */

/* ------------------------------------------------------------------------------------------
 * small helpers [kernels.cuh]
 * ---------------------------------------------------------------------------------------- */
int64_t swo_sub2ind(int64_t x, int64_t y, int64_t z, int64_t nx, int64_t ny, int64_t nz)
{
    (void)nx;
    return x * nz * ny + y * nz + z; /* [kernels.cuh:53-60] row-major, x slowest */
}

void swo_xrot(float s, float c, const float *m0, float *m1)
{ /* [kernels.cuh:82-87] */
    m1[0] = m0[0];
    m1[1] = c * m0[1] - s * m0[2];
    m1[2] = s * m0[1] + c * m0[2];
}

void swo_yrot(float s, float c, const float *m0, float *m1)
{ /* [kernels.cuh:108-113] */
    m1[0] = c * m0[0] + s * m0[2];
    m1[1] = m0[1];
    m1[2] = -s * m0[0] + c * m0[2];
}

void swo_zrot(float s, float c, const float *m0, float *m1)
{ /* [kernels.cuh:134-139] */
    m1[0] = c * m0[0] - s * m0[1];
    m1[1] = s * m0[0] + c * m0[1];
    m1[2] = m0[2];
}

static void zrot_deg(float theta, const float *m0, float *m1)
{ /* [kernels.cuh:147-152] theta(float)*DEG2RAD(double) -> float argument of sinf/cosf */
    float a = (float)(theta * SWO_DEG2RAD);
    swo_zrot(sinf(a), cosf(a), m0, m1);
}

void swo_relax(float e1, float e2, const float *m0, float *m1)
{ /* [kernels.cuh:214-219] the z line is evaluated in double (literal 1.) */
    float z = m0[2];
    m1[0] = m0[0] * e2;
    m1[1] = m0[1] * e2;
    m1[2] = (float)(1. + (double)e1 * ((double)z - 1.));
}

void swo_xrot_withphase(float s, float c, float ph, const float *m0, float *m1)
{ /* [kernels.cuh:160-195] exact fast paths, else Rz(ph) Rx(theta) Rz(-ph) */
    if (ph == 0.0f) { swo_xrot(s, c, m0, m1); return; }
    if (ph == 180.0) { swo_xrot(-s, c, m0, m1); return; }
    if (ph == 90.0) { swo_yrot(s, c, m0, m1); return; }
    if (ph == -90.0 || ph == 270.0) { swo_yrot(-s, c, m0, m1); return; }
    float t[3];
    float a = (float)(ph * SWO_DEG2RAD);
    float sp = sinf(a), cp = cosf(a);
    swo_zrot(-sp, cp, m0, m1);
    swo_xrot(s, c, m1, t);
    swo_zrot(sp, cp, t, m1);
}

static void xrot_withphase_deg(float theta, float ph, const float *m0, float *m1)
{ /* [kernels.cuh:203-206] double-precision sin/cos of the flip angle, narrowed to float */
    swo_xrot_withphase((float)sin(theta * SWO_DEG2RAD), (float)cos(theta * SWO_DEG2RAD), ph, m0, m1);
}

/* [kernels.cu:45-52].  exp_is_double: which overload `exp(float)` resolves to in the build being
 * restated — g++ host build sees only ::exp(double); nvcc sees the CUDA float overload (expf). */
static void dephase_relax(const float *m0, float *m1, float acc_phase, float T1, float T2, float dt_s, int exp_is_double)
{
    zrot_deg(acc_phase, m0, m1);
    if (T1 >= 0 && T2 >= 0) {
        float e1, e2;
        if (exp_is_double) {
            e1 = (float)exp((double)(-dt_s / T1));
            e2 = (float)exp((double)(-dt_s / T2));
        } else {
            e1 = expf(-dt_s / T1);
            e2 = expf(-dt_s / T2);
        }
        swo_relax(e1, e2, m1, m1);
    }
}

/* ------------------------------------------------------------------------------------------
 * libstdc++ flavour: mt19937 + generate_canonical<float,24> + polar normal
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t x[624]; int p; } mt_t;

static void mt_seed(mt_t *g, uint64_t sd)
{ /* mersenne_twister_engine::seed: value mod 2^32, f = 1812433253 */
    g->x[0] = (uint32_t)sd;
    for (int i = 1; i < 624; i++) g->x[i] = 1812433253u * (g->x[i - 1] ^ (g->x[i - 1] >> 30)) + (uint32_t)i;
    g->p = 624;
}

static void mt_refill(mt_t *g)
{
    uint32_t *x = g->x;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (x[k] & 0x80000000u) | (x[(k + 1) % 624] & 0x7fffffffu);
        x[k] = x[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->p = 0;
}

static uint32_t mt_next(mt_t *g)
{
    if (g->p >= 624) mt_refill(g);
    uint32_t z = g->x[g->p++];
    z ^= (z >> 11);
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= (z >> 18);
    return z;
}

static void mt_discard(mt_t *g, uint64_t z)
{ /* discard(z) drops z outputs; tempering has no state so only the index/refill matters */
    while (z > (uint64_t)(624 - g->p)) {
        z -= (uint64_t)(624 - g->p);
        mt_refill(g);
    }
    g->p += (int)z;
}

static float mt_canonical(mt_t *g)
{ /* generate_canonical<float,24> with a 2^32-range engine: one draw [random.tcc:3354-3380] */
    float sum = (float)mt_next(g); /* uint -> float rounds to nearest, may hit 2^32 */
    float ret = sum / 4294967296.0f;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret;
}

typedef struct { mt_t g; int saved_ok; float saved; } mt_normal_t;

static float mt_normal(mt_normal_t *n)
{ /* Marsaglia polar [random.tcc:1809-1844], mean 0 stddev 1 */
    float ret;
    if (n->saved_ok) {
        n->saved_ok = 0;
        ret = n->saved;
    } else {
        float x, y, r2;
        do {
            x = (float)(2.0f * mt_canonical(&n->g) - 1.0);
            y = (float)(2.0f * mt_canonical(&n->g) - 1.0);
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        float mult = sqrtf(-2 * logf(r2) / r2);
        n->saved = x * mult;
        n->saved_ok = 1;
        ret = y * mult;
    }
    return ret * 1.0f + 0.0f;
}

/* ------------------------------------------------------------------------------------------
 * Thrust flavour: minstd_rand + erfcinv normal + uniform
 * ---------------------------------------------------------------------------------------- */
#define MINSTD_A 48271u
#define MINSTD_M 2147483647u

static uint32_t minstd_seed(uint64_t sd)
{ /* ctor takes result_type = uint32: truncation, then mod m, 0 -> 1
     [thrust/random/detail/linear_congruential_engine.inl:45-55] */
    uint32_t s = (uint32_t)sd % MINSTD_M;
    return s == 0 ? 1u : s;
}

static uint32_t minstd_next(uint32_t *x)
{
    *x = (uint32_t)(((uint64_t)*x * MINSTD_A) % MINSTD_M);
    return *x;
}

static void minstd_discard(uint32_t *x, uint64_t z)
{ /* x <- a^z x mod m [thrust/random/detail/linear_congruential_engine_discard.h] */
    uint64_t mult = MINSTD_A, acc = 1;
    while (z > 0) {
        if (z & 1) acc = (acc * mult) % MINSTD_M;
        z >>= 1;
        mult = (mult * mult) % MINSTD_M;
    }
    *x = (uint32_t)((acc * *x) % MINSTD_M);
}

/* Double-precision inverse complementary error function on (0,2).
 * Start: M. Giles, "Approximating the erfinv function" (single-precision polynomial, evaluated on
 * w = -log(x(2-x)) so that small x does not cancel); polish: Newton on libm erfc. */
double swo_erfcinv(double x)
{
    if (!(x > 0.0)) return x == 0.0 ? INFINITY : NAN;
    if (!(x < 2.0)) return x == 2.0 ? -INFINITY : NAN;
    double z = 1.0 - x;
    double w = -log(x * (2.0 - x));
    double p;
    if (w < 5.0) {
        w -= 2.5;
        p = 2.81022636e-08;
        p = 3.43273939e-07 + p * w;
        p = -3.5233877e-06 + p * w;
        p = -4.39150654e-06 + p * w;
        p = 0.00021858087 + p * w;
        p = -0.00125372503 + p * w;
        p = -0.00417768164 + p * w;
        p = 0.246640727 + p * w;
        p = 1.50140941 + p * w;
    } else {
        w = sqrt(w) - 3.0;
        p = -0.000200214257;
        p = 0.000100950558 + p * w;
        p = 0.00134934322 + p * w;
        p = -0.00367342844 + p * w;
        p = 0.00573950773 + p * w;
        p = -0.0076224613 + p * w;
        p = 0.00943887047 + p * w;
        p = 1.00167406 + p * w;
        p = 2.83297682 + p * w;
    }
    double y = p * z;
    const double two_over_sqrtpi = 1.12837916709551257390;
    for (int it = 0; it < 60; it++) {
        double f = erfc(y) - x;
        double d = f / (two_over_sqrtpi * exp(-y * y)); /* y_new = y + f/|f'| */
        y += d;
        if (fabs(d) <= 2e-16 * fabs(y) || d == 0.0) break;
    }
    return y;
}

static float minstd_normal(uint32_t *x)
{ /* normal_distribution_nvcc::sample, mean 0 stddev 1, HOST arithmetic
     [thrust/random/detail/normal_distribution_base.h:52-80] */
    const uint32_t range = 2147483646u - 1u;                    /* max - min */
    const float S1 = (float)(1. / (double)range), S2 = S1 / 2;
    float S3 = (float)(-1.4142135623730950488016887242097);
    uint32_t u = minstd_next(x) - 1u;
    if (u > range / 2) {
        u = range - u;
        S3 = -S3;
    }
    float p = (float)u * S1 + S2;
    float e = (float)swo_erfcinv((double)(2 * p)); /* host erfcinvf [crt/math_functions.hpp:3367] */
    return 0.0f + 1.0f * S3 * e;
}

static float minstd_uniform(uint32_t *x)
{ /* [thrust/random/detail/uniform_real_distribution.inl:61-75] with a=0, b=1 */
    float r = (float)(minstd_next(x) - 1u);
    r /= (1.0f + (float)(2147483646u - 1u));
    return r * (1.0f - 0.0f) + 0.0f;
}

/* exposed streams for unit tests */
void swo_minstd_normals(uint64_t sps, uint32_t n, float *out)
{
    uint32_t x = minstd_seed(sps);
    minstd_discard(&x, sps);
    for (uint32_t i = 0; i < n; i++) out[i] = minstd_normal(&x);
}
void swo_minstd_uniforms(uint64_t sps, uint32_t n, float *out)
{
    uint32_t x = minstd_seed(sps);
    minstd_discard(&x, sps);
    for (uint32_t i = 0; i < n; i++) out[i] = minstd_uniform(&x);
}
void swo_mt_normals(uint64_t sps, uint32_t n, float *out)
{
    mt_normal_t g;
    mt_seed(&g.g, sps);
    g.saved_ok = 0;
    mt_discard(&g.g, sps);
    for (uint32_t i = 0; i < n; i++) out[i] = mt_normal(&g);
}
void swo_mt_uniforms(uint64_t sps, uint32_t n, float *out)
{
    mt_t g;
    mt_seed(&g, sps);
    mt_discard(&g, sps);
    for (uint32_t i = 0; i < n; i++) out[i] = mt_canonical(&g);
}

/* ------------------------------------------------------------------------------------------
 * parameters::prepare  [simulation_parameters.cuh:227-245]
 * ---------------------------------------------------------------------------------------- */
uint32_t swo_n_timepoints(const swo_case *c) { return (uint32_t)(c->TR_us / c->timestep_us); }

int32_t swo_n_dummy_scan(const swo_case *c)
{
    if (c->n_dummy_scan >= 0) return c->n_dummy_scan;
    return (int32_t)(5.0 * c->T1_ms[0] / (float)(c->TR_us * 1e-3));
}

double swo_step_sigma(double D, int32_t timestep_us) { return 1e-3 * sqrt(2. * D * timestep_us); }

float swo_tesla_to_deg_per_step(float B0, int32_t timestep_us)
{ /* [monte_carlo.cu:241] float*int -> float, then double chain, narrowed to float */
    return (float)((double)(B0 * (float)timestep_us) * 1e-6 * SWO_GAMMA * SWO_RAD2DEG);
}

typedef struct prep {
    const swo_case *c;
    double   fov[3];
    float    cs, sn; /* cos / sin of RF_FA[0] */
    float    lin_pc, quad_pc;
    uint32_t n_timepoints;
    int32_t  n_dummy_scan;
    int64_t  matrix_length;
    size_t   trj;
    double  *sigma;             /* [n_substrate] per-axis step sigma, metres */
    const float *gx, *gy, *gz;  /* possibly scaled gradient tables            */
    int      flavour;
} prep;

/* ------------------------------------------------------------------------------------------
 * the time loop for one spin  [kernels.cu:72-233]
 * ---------------------------------------------------------------------------------------- */
typedef struct rng {
    int flavour;
    mt_normal_t r; /* gen_r + normal distribution state */
    mt_t u;        /* gen_u                              */
    uint32_t xr, xu;
} rng;

static void rng_init(rng *g, int flavour, uint64_t seed_plus_spin)
{ /* [kernels.cu:76-88] both engines get the same seed and the same discard */
    g->flavour = flavour;
    if (flavour == SWO_RNG_MT19937) {
        mt_seed(&g->r.g, seed_plus_spin);
        g->r.saved_ok = 0;
        mt_discard(&g->r.g, seed_plus_spin);
        g->u = g->r.g;
    } else {
        g->xr = minstd_seed(seed_plus_spin);
        minstd_discard(&g->xr, seed_plus_spin);
        g->xu = g->xr;
    }
}
static float rng_normal(rng *g) { return g->flavour == SWO_RNG_MT19937 ? mt_normal(&g->r) : minstd_normal(&g->xr); }
static float rng_uniform(rng *g) { return g->flavour == SWO_RNG_MT19937 ? mt_canonical(&g->u) : minstd_uniform(&g->xu); }

static void sim_spin(const prep *P, const float *fieldmap, const uint8_t *mask, const float *M0, const float *XYZ0,
                     float *M1, float *XYZ1, uint8_t *T, uint32_t spin_no, swo_stats *st)
{
    const swo_case *c = P->c;
    const int exp_dbl = (P->flavour == SWO_RNG_MT19937);
    float *xyz1 = XYZ1 + 3 * (size_t)spin_no * P->trj; /* [kernels.cu:75] */
    rng g;
    rng_init(&g, P->flavour, c->seed + spin_no);

    uint32_t itr = 0;
    float field = 0.f, T1 = 0.f, T2 = 0.f, rf_phase = c->RF_PH_deg[0], time_elapsed = 0.f; /* [kernels.cu:91] */
    float m0[3], m1[3];
    double xyz_old[3], xyz_new[3], scale2grid[3];
    const size_t shift = 3 * (size_t)spin_no;
    for (int i = 0; i < 3; i++) { /* [kernels.cu:95-99] */
        xyz_old[i] = xyz_new[i] = xyz1[i] = XYZ0[shift + i];
        m0[i] = M0[shift + i];
        scale2grid[i] = (double)c->phantom_size[i] / P->fov[i];
    }
    const int64_t nx = (int64_t)c->phantom_size[0], ny = (int64_t)c->phantom_size[1], nz = (int64_t)c->phantom_size[2];
    uint8_t ts, ts_old; /* [kernels.cu:101-104] */
    int64_t indx = swo_sub2ind((int64_t)(xyz1[0] * scale2grid[0]), (int64_t)(xyz1[1] * scale2grid[1]),
                               (int64_t)(xyz1[2] * scale2grid[2]), nx, ny, nz);
    ts = ts_old = mask[indx];
    double sigma = P->sigma[ts_old];

    const uint32_t n_dummy = (uint32_t)P->n_dummy_scan;
    for (uint32_t dummy_scan = 0; dummy_scan < n_dummy + 1; dummy_scan++) { /* [kernels.cu:107] */
        int is_lastscan = (dummy_scan == n_dummy);
        /* [kernels.cu:110-114] float + float, then + double, narrowed; wrap into [0,360] */
        float new_rf_phase = (float)((double)(rf_phase + (float)dummy_scan * P->lin_pc) +
                                     (double)(uint32_t)(dummy_scan * (dummy_scan + 1)) / 2.0 * (double)P->quad_pc);
        while (new_rf_phase > 360.0) new_rf_phase = (float)(new_rf_phase - 360.0);
        while (new_rf_phase < 0) new_rf_phase = (float)(new_rf_phase + 360.0);

        swo_xrot_withphase(P->sn, P->cs, new_rf_phase, m0, m1); /* [kernels.cu:117] */
        for (int i = 0; i < 3; i++) m0[i] = m1[i];

        int64_t ind = 0, ind_old = P->matrix_length + 1; /* [kernels.cu:123-126] */
        uint32_t tp = 0, tp_old = 0;
        uint16_t cur_rf = 1, cur_te = 0, cnt_deph = 0, cnt_grad = 0;
        float acc_phase = 0.f;

        while (tp < P->n_timepoints) { /* [kernels.cu:128] */
            for (int i = 0; i < 3 && sigma != 0.; i++) { /* [kernels.cu:130-137] */
                double rnd = (double)rng_normal(&g) * sigma;
                xyz_new[i] = xyz_old[i] + rnd;
                if (xyz_new[i] < 0)
                    xyz_new[i] += (c->cross_fov ? P->fov[i] : 2 * fabs(rnd));
                else if (xyz_new[i] >= P->fov[i])
                    xyz_new[i] -= (c->cross_fov ? P->fov[i] : 2 * fabs(rnd));
            }
            /* [kernels.cu:140-147] double -> int64 truncation, range guard */
            ind = swo_sub2ind((int64_t)(xyz_new[0] * scale2grid[0]), (int64_t)(xyz_new[1] * scale2grid[1]),
                              (int64_t)(xyz_new[2] * scale2grid[2]), nx, ny, nz);
            if (ind >= P->matrix_length || ind < 0) {
                st->lost++;
                return;
            }
            if (ind != ind_old) { /* [kernels.cu:150-170] */
                st->mask_gathers++;
                ts = mask[ind];
                if (ts != ts_old) {
                    if (rng_uniform(&g) >= c->pXY[ts_old * c->n_substrate + ts]) {
                        st->rejects++;
                        if (itr++ > c->max_iterations) {
                            st->lost++;
                            return;
                        }
                        continue;
                    }
                    ts_old = ts;
                }
                ind_old = ind;
                st->field_gathers++;
                field = fieldmap ? fieldmap[ind] : 0.f;
                T1 = (float)(c->T1_ms[ts_old] * 1e-3);
                T2 = (float)(c->T2_ms[ts_old] * 1e-3);
                sigma = P->sigma[ts_old];
            }
            acc_phase += field; /* [kernels.cu:171-172] */
            itr = 0;
            st->steps++;

            if (cnt_deph < c->n_dephasing && (uint32_t)c->dephasing_tp[cnt_deph] == tp) { /* [kernels.cu:175-178] */
                acc_phase += (float)spin_no * c->dephasing_deg[cnt_deph] / (float)c->n_spins;
                cnt_deph++;
            }
            if (cnt_grad < c->n_gradient && (uint32_t)c->gradient_tp[cnt_grad] == tp) { /* [kernels.cu:181-187] */
                const float Gx = P->gx[cnt_grad], Gy = P->gy[cnt_grad], Gz = P->gz[cnt_grad];
                acc_phase = (float)((double)acc_phase +
                                    (Gx * xyz_new[0] + Gy * xyz_new[1] + Gz * xyz_new[2]) * 1e-3 * c->timestep_us * 1e-6 *
                                        SWO_GAMMA * SWO_RAD2DEG);
                cnt_grad++;
            }
            if (cur_rf < c->n_RF && (uint32_t)c->RF_tp[cur_rf] == tp) { /* [kernels.cu:190-199] */
                time_elapsed = (float)((uint32_t)((tp - tp_old) * (uint32_t)c->timestep_us) * 1e-6);
                dephase_relax(m0, m1, acc_phase, T1, T2, time_elapsed, exp_dbl);
                xrot_withphase_deg(c->RF_FA_deg[cur_rf], c->RF_PH_deg[cur_rf], m1, m0);
                acc_phase = 0;
                tp_old = tp;
                cur_rf++;
            }
            if (is_lastscan && cur_te < c->n_TE && (uint32_t)c->TE_tp[cur_te] == tp) { /* [kernels.cu:202-215] */
                time_elapsed = (float)((uint32_t)((tp - tp_old) * (uint32_t)c->timestep_us) * 1e-6);
                dephase_relax(m0, m1, acc_phase, T1, T2, time_elapsed, exp_dbl);
                size_t sh = 3 * (size_t)c->n_TE * spin_no + 3 * (size_t)cur_te;
                for (int i = 0; i < 3; i++) M1[sh + i] = m0[i] = m1[i];
                T[(size_t)spin_no * c->n_TE + cur_te] = ts_old;
                acc_phase = 0;
                tp_old = tp;
                cur_te++;
            }
            if (c->record_trajectory && (tp != 0 || dummy_scan != 0)) xyz1 += 3; /* [kernels.cu:218-221] */
            for (int i = 0; i < 3; i++) {
                xyz_old[i] = xyz_new[i];
                xyz1[i] = (float)xyz_new[i];
            }
            tp++;
        }
        /* [kernels.cu:226-231] */
        time_elapsed = (float)((uint32_t)((tp - tp_old) * (uint32_t)c->timestep_us) * 1e-6);
        dephase_relax(m0, m1, acc_phase, T1, T2, time_elapsed, exp_dbl);
        for (int i = 0; i < 3; i++) m0[i] = m1[i];
    }
}

/* ------------------------------------------------------------------------------------------
 * host driver  [monte_carlo.cu:236-244, 264-337]
 * ---------------------------------------------------------------------------------------- */
typedef struct job {
    const prep *P;
    const float *fieldmap;
    const uint8_t *mask;
    const float *M0, *XYZ0;
    float *M1, *XYZ1;
    uint8_t *T;
    uint32_t begin, end;
    volatile uint32_t *next;
    swo_stats st;
} job;

static void *worker(void *arg)
{
    job *j = (job *)arg;
    const uint32_t chunk = 64;
    for (;;) {
        uint32_t b = __atomic_fetch_add(j->next, chunk, __ATOMIC_RELAXED);
        if (b >= j->end) break;
        uint32_t e = b + chunk < j->end ? b + chunk : j->end;
        for (uint32_t s = b; s < e; s++) sim_spin(j->P, j->fieldmap, j->mask, j->M0, j->XYZ0, j->M1, j->XYZ1, j->T, s, &j->st);
    }
    return NULL;
}

int swo_run(const swo_case *c, const float *fieldmap_T, const uint8_t *mask, const float *XYZ0, const float *M0,
            float *M1, float *XYZ1, uint8_t *T, uint32_t spin_begin, uint32_t spin_end, int flavour, int n_threads,
            swo_stats *stats, double *seconds)
{
    if (!c || !mask || !XYZ0 || !M0 || !M1 || !XYZ1 || !T) return -1;
    if (c->n_RF < 1 || c->n_substrate < 1 || c->timestep_us <= 0 || c->seed == 0) return -2;
    if (spin_end > c->n_spins) spin_end = c->n_spins;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;

    const size_t V = (size_t)c->phantom_size[0] * c->phantom_size[1] * c->phantom_size[2];
    const size_t S = c->n_spins, nTE = c->n_TE, nG = c->n_gradient;
    prep P;
    memset(&P, 0, sizeof P);
    P.c = c;
    P.flavour = flavour;
    { /* parameters::prepare [simulation_parameters.cuh:229-230]: cosf/sinf(float(FA*DEG2RAD)) */
        float a = (float)(c->RF_FA_deg[0] * SWO_DEG2RAD);
        P.cs = cosf(a);
        P.sn = sinf(a);
    }
    P.n_timepoints = swo_n_timepoints(c);
    P.n_dummy_scan = swo_n_dummy_scan(c);
    P.matrix_length = (int64_t)V;
    P.trj = c->record_trajectory ? (size_t)P.n_timepoints * (size_t)(P.n_dummy_scan + 1) : 1;
    P.sigma = (double *)malloc(sizeof(double) * c->n_substrate);
    for (uint32_t i = 0; i < c->n_substrate; i++) P.sigma[i] = swo_step_sigma(c->diffusivity[i], c->timestep_us);
    P.lin_pc = c->linear_phase_cycling;
    P.quad_pc = c->quadratic_phase_cycling;
    for (int i = 0; i < 3; i++) P.fov[i] = (float)c->fov[i]; /* fov is held as float [monte_carlo.cuh:37] */

    float *fm = NULL; /* [monte_carlo.cu:241-244] */
    if (fieldmap_T) {
        float k = swo_tesla_to_deg_per_step(c->B0, c->timestep_us);
        fm = (float *)malloc(sizeof(float) * V);
        for (size_t i = 0; i < V; i++) fm[i] = fieldmap_T[i] * k;
    }
    float *xyz0s = (float *)malloc(sizeof(float) * 3 * S);
    memcpy(xyz0s, XYZ0, sizeof(float) * 3 * S);
    float *g = (float *)malloc(sizeof(float) * 3 * (nG ? nG : 1));
    for (size_t i = 0; i < nG; i++) { g[i] = c->gradX_mTm[i]; g[nG + i] = c->gradY_mTm[i]; g[2 * nG + i] = c->gradZ_mTm[i]; }
    P.gx = g; P.gy = g + nG; P.gz = g + 2 * nG;

    swo_stats tot;
    memset(&tot, 0, sizeof tot);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t k = 0; k < c->n_scales; k++) {
        const float scale = c->scales[k];
        if (c->scale_type == SWO_SCALE_FOV) { /* [monte_carlo.cu:277-285] float*float */
            for (size_t i = 0; i < 3 * S; i++) xyz0s[i] = XYZ0[i] * scale;
            for (int i = 0; i < 3; i++) P.fov[i] = scale * (float)c->fov[i];
        } else if (c->scale_type == SWO_SCALE_GRADIENT) { /* [monte_carlo.cu:287-301] */
            for (size_t i = 0; i < nG; i++) {
                g[i] = c->gradX_mTm[i] * scale;
                g[nG + i] = c->gradY_mTm[i] * scale;
                g[2 * nG + i] = c->gradZ_mTm[i] * scale;
            }
        } else if (c->scale_type == SWO_SCALE_PHASE_CYCLING) { /* [monte_carlo.cu:302-305] */
            P.lin_pc = c->linear_phase_cycling * scale;
            P.quad_pc = c->quadratic_phase_cycling;
        }
        volatile uint32_t next = spin_begin;
        job *jobs = (job *)calloc((size_t)n_threads, sizeof(job));
        pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
        for (int i = 0; i < n_threads; i++) {
            jobs[i].P = &P;
            jobs[i].fieldmap = fm;
            jobs[i].mask = mask;
            jobs[i].M0 = M0;
            jobs[i].XYZ0 = xyz0s;
            jobs[i].M1 = M1 + 3 * nTE * S * k; /* [monte_carlo.cu:318-320] */
            jobs[i].XYZ1 = XYZ1 + 3 * S * P.trj * k;
            jobs[i].T = T + nTE * S * k;
            jobs[i].begin = spin_begin;
            jobs[i].end = spin_end;
            jobs[i].next = &next;
        }
        for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, worker, &jobs[i]);
        worker(&jobs[0]);
        for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
        for (int i = 0; i < n_threads; i++) {
            tot.steps += jobs[i].st.steps;
            tot.mask_gathers += jobs[i].st.mask_gathers;
            tot.field_gathers += jobs[i].st.field_gathers;
            tot.rejects += jobs[i].st.rejects;
            tot.lost += jobs[i].st.lost;
        }
        free(jobs);
        free(th);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (stats) *stats = tot;
    free(g);
    free(xyz0s);
    free(fm);
    free(P.sigma);
    return 0;
}

/* [monte_carlo.cu:142-151] std::mt19937(seed) + uniform_real_distribution<float>(0.01 fov, 0.99 fov),
 * drawn x,y,z interleaved; each draw = canonical*(b-a)+a in float. */
void swo_init_positions(uint64_t seed, const float fov[3], uint32_t n_spins, float *XYZ0)
{
    mt_t g;
    mt_seed(&g, seed);
    float a[3], b[3];
    for (int i = 0; i < 3; i++) {
        a[i] = (float)(0.01 * fov[i]);
        b[i] = (float)(0.99 * fov[i]);
    }
    for (size_t s = 0; s < n_spins; s++)
        for (int i = 0; i < 3; i++) XYZ0[3 * s + i] = mt_canonical(&g) * (b[i] - a[i]) + a[i];
}
