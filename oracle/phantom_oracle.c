/* oracle/phantom_oracle.c — TEST INFRASTRUCTURE (never linked into, imported by, or called from the product path).
 *
 * Plain-C restatement of the reference's `spinwalk phantom -c | -s | -t` generators, single-threaded:
 *   voxel-centre grid        src/phantom/phantom_base.cpp:107-143
 *   cylinders: placement     src/phantom/phantom_cylinder.cpp:22-56 (overlap), :85-130 (loop), :133-181 (volume)
 *              voxel fill    src/phantom/phantom_cylinder.cpp:184-275
 *   spheres:   placement     src/phantom/phantom_sphere.cpp:22-55 (overlap), :79-119 (loop)
 *              voxel fill    src/phantom/phantom_sphere.cpp:121-198
 *   two pools                src/phantom/phantom_twopools.cpp:40-63
 * The reference computes with float variables but promotes through the double constant M_PI in places; every expression
 * below keeps the reference's types and evaluation order, and the file is compiled with -ffp-contract=off, so that the
 * output is the bit pattern a baseline x86-64 build of the reference produces.
 *
 * Third-party arithmetic: libstdc++ <random> (GCC 13.3 here; reference Docker: GCC 11): std::mt19937 (cylinders),
 * std::minstd_rand (spheres), std::uniform_real_distribution<float>(0,1) = generate_canonical<float,24> with one engine
 * draw (/usr/include/c++/13/bits/random.tcc:3354-3380).
 *
 * PINNED: bit-exact (shape list, mask, fieldmap, volume fraction) against the UNMODIFIED reference sources compiled by
 * oracle/Makefile into oracle/_ref/libswref_gen.so (tests/test_phantom_oracle.py) and against tests/golden/phantom_*.npz.
 * The OpenMP build of the reference visits shapes in a thread-dependent order inside check_*_overlap when radii are random
 * (phantom_cylinder.cpp:26, a race on `radius`); this restatement follows the serial order (a build without -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "phantom_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- engines (libstdc++) ---- */
typedef struct { uint32_t x[624]; int p; } mt_t;
static void mt_seed(mt_t *g, uint64_t sd)
{
    g->x[0] = (uint32_t)sd;
    for (int i = 1; i < 624; i++) g->x[i] = 1812433253u * (g->x[i - 1] ^ (g->x[i - 1] >> 30)) + (uint32_t)i;
    g->p = 624;
}
static uint32_t mt_next(mt_t *g)
{
    if (g->p >= 624) {
        uint32_t *x = g->x;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (x[k] & 0x80000000u) | (x[(k + 1) % 624] & 0x7fffffffu);
            x[k] = x[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->p = 0;
    }
    uint32_t z = g->x[g->p++];
    z ^= (z >> 11);
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= (z >> 18);
    return z;
}
/* std::minstd_rand = linear_congruential_engine<uint_fast32_t, 48271, 0, 2147483647>; seed s: s mod m, 0 -> 1 */
typedef struct { uint64_t x; } lcg_t;
static void lcg_seed(lcg_t *g, uint64_t sd)
{
    g->x = sd % 2147483647ull;
    if (g->x == 0) g->x = 1;
}
static uint32_t lcg_next(lcg_t *g)
{
    g->x = (g->x * 48271ull) % 2147483647ull;
    return (uint32_t)g->x;
}

typedef struct { int kind; mt_t mt; lcg_t lcg; } urng_t; /* kind 0 = mt19937, 1 = minstd_rand */
/* uniform_real_distribution<float>(0,1)(gen) -> generate_canonical<float,24>: k = 1 draw for both engines;
 * ret = float(draw - min) / float(range) with range = 2^32 (mt19937) or 2^31-2 -> 2147483648.0f after rounding to float */
static float canonical(urng_t *g)
{
    float sum, tmp;
    if (g->kind == 0) { sum = (float)mt_next(&g->mt); tmp = 4294967296.0f; }
    else { sum = (float)(lcg_next(&g->lcg) - 1u); tmp = (float)2147483646.0L; }
    float ret = sum / tmp;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret;
}

/* ---- phantom_base::create_grid [phantom_base.cpp:122-129]: voxel centres, double arithmetic stored as float ---- */
static void grid_base(float fov, size_t resolution, float *g)
{
    const double start = fov / resolution / 2.0; /* float / size_t is a float division, then / 2.0 in double */
    const double end = fov - fov / resolution / 2.0;
    const double step = (end - start) / (resolution - 1.0);
    for (size_t i = 0; i < resolution; i++) g[i] = (float)(start + i * step);
}

static int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }

/* ---- cylinders ---- */
/* [phantom_cylinder.cpp:22-56], serial order */
static int cyl_overlap(const float *pts, const float *radii, size_t n, const float *cyl_pnt, float *radius, int is_random_radius)
{
    for (size_t c = 0; c < n; c++) {
        float p0 = cyl_pnt[0] - pts[3 * c], p1 = cyl_pnt[1] - pts[3 * c + 1];
        float distance = sqrtf(p0 * p0 + p1 * p1);
        if (distance <= radii[c] || distance <= *radius) return 1;
        else if (distance < radii[c] + *radius) {
            if (!is_random_radius) return 1;
            *radius = distance - radii[c];
        }
    }
    return 0;
}

/* [phantom_cylinder.cpp:133-181] */
static float cyl_volume(const float *g, float fov, size_t resolution, const float *cyl_pnt, float cyl_rad)
{
    int intersect = 0;
    for (int i = 0; i < 2; i++)
        if (cyl_pnt[i] + cyl_rad < 0 || cyl_pnt[i] - cyl_rad > fov) return -1.f;
    for (int i = 0; i < 2; i++)
        if (cyl_pnt[i] < cyl_rad - 1.5 || cyl_pnt[i] > fov - cyl_rad + 1.5) intersect = 1;
    if (!intersect) return (float)(M_PI * cyl_rad * cyl_rad * fov);
    float v_size = fov / resolution;
    float cyl_rad2 = cyl_rad * cyl_rad;
    int32_t vox[2] = {(int32_t)(cyl_pnt[0] / v_size), (int32_t)(cyl_pnt[1] / v_size)};
    int32_t rad_vox = (int32_t)(ceilf(cyl_rad / fov * resolution) + 1);
    int32_t x_min = imax(0, vox[0] - rad_vox), x_max = imin((int32_t)resolution, vox[0] + rad_vox + 2);
    int32_t y_min = imax(0, vox[1] - rad_vox), y_max = imin((int32_t)resolution, vox[1] + rad_vox + 2);
    int32_t counter = 0;
    for (int32_t py = y_min; py < y_max; py++)
        for (int32_t px = x_min; px < x_max; px++) {
            float p0 = g[px] - cyl_pnt[0], p1 = g[py] - cyl_pnt[1];
            float distance2 = p0 * p0 + p1 * p1;
            if (distance2 <= cyl_rad2) counter++;
        }
    counter *= (int32_t)resolution; /* the reference also loops pz over [0, resolution): the test does not depend on z */
    return counter * v_size * v_size * v_size;
}

/* [phantom_cylinder.cpp:85-130] */
static int cyl_place(const swo_phantom_spec *s, const float *g, float **pts_out, float **radii_out, uint32_t *n_out)
{
    const float fov = s->fov_um, radius = s->radius_um, vf = s->volume_fraction;
    if (2 * radius >= fov) return 1;
    const int is_random_radius = radius < 0;
    const float max_radius = radius > 0 ? radius : -radius;
    size_t cap = 1024, n = 0;
    float *pts = malloc(cap * 3 * sizeof(float)), *radii = malloc(cap * sizeof(float));
    urng_t gen;
    gen.kind = 0;
    mt_seed(&gen.mt, (uint64_t)s->seed);
    float cyl_pnt[3], cyl_rad, vol_cyl = 0, vol_cyl_total = 0, vol_tol = fov * fov * fov;
    int32_t progress = 0;
    while (progress < 100) {
        cyl_rad = is_random_radius ? canonical(&gen) * max_radius : max_radius;
        for (int i = 0; i < 3; i++) cyl_pnt[i] = canonical(&gen) * (fov + 2 * cyl_rad) - cyl_rad;
        if (cyl_overlap(pts, radii, n, cyl_pnt, &cyl_rad, is_random_radius)) continue;
        vol_cyl = cyl_volume(g, fov, s->resolution, cyl_pnt, cyl_rad);
        if (100 * (vol_cyl + vol_cyl_total) / vol_tol > 1.02 * vf || vol_cyl < 0) continue;
        vol_cyl_total += vol_cyl;
        progress = (int32_t)(100 * (100. * vol_cyl_total / vol_tol / vf));
        if (n == cap) {
            cap *= 2;
            pts = realloc(pts, cap * 3 * sizeof(float));
            radii = realloc(radii, cap * sizeof(float));
        }
        memcpy(pts + 3 * n, cyl_pnt, 3 * sizeof(float));
        radii[n++] = cyl_rad;
    }
    *pts_out = pts;
    *radii_out = radii;
    *n_out = (uint32_t)n;
    return 0;
}

/* roty [phantom_base.h:133-142], T = float.  The unqualified sin/cos there resolve to the double functions of <math.h>
 * (the reference build imports sincos, not sincosf): the float argument is widened and the double result narrowed. */
static void roty_f(float theta, const float *m0, float *m1)
{
    float deg2rad = (float)0.0174532925199433;
    float sn = (float)sin((double)(theta * deg2rad)), cs = (float)cos((double)(theta * deg2rad));
    m1[0] = cs * m0[0] + sn * m0[2];
    m1[1] = m0[1];
    m1[2] = -sn * m0[0] + cs * m0[2];
}

/* [phantom_cylinder.cpp:184-275] */
static void cyl_fill(const swo_phantom_spec *s, const float *g, const float *pts, const float *radii, uint32_t n, uint8_t *mask, float *fieldmap,
                     int32_t zlo, int32_t zhi)
{ /* outputs hold the z window [zlo, zhi) only: [res][res][zhi-zlo] (the whole volume for 0, res) */
    const size_t res1 = s->resolution, nzw = (size_t)(zhi - zlo);
    const int calc = fieldmap != NULL;
    const float fov = s->fov_um, Y = s->Y, dChi = s->dchi;
    float v_size = fov / res1;
    float B0_orig[3] = {0.f, 0.f, 1.f}, B0[3];
    roty_f(s->orientation_deg, B0_orig, B0);
    float B0_prj[3] = {B0[0], B0[1], 0.0f};
    { /* normalize(B0_prj): n = sqrt(float) -> float; n == 0 leaves it */
        float nn = sqrtf(B0_prj[0] * B0_prj[0] + B0_prj[1] * B0_prj[1] + B0_prj[2] * B0_prj[2]);
        if (nn != 0) { B0_prj[0] /= nn; B0_prj[1] /= nn; B0_prj[2] /= nn; }
    }
    float theta_c = (float)cos(s->orientation_deg * M_PI / 180);
    float theta_c2 = theta_c * theta_c;
    float theta_s2 = (float)(1. - theta_c2);
    for (uint32_t c = 0; c < n; c++) {
        const float *cyl_pnt = pts + 3 * c;
        float cyl_rad = radii[c], cyl_rad2 = cyl_rad * cyl_rad;
        int32_t rad_vox = (int32_t)(ceilf(cyl_rad / v_size) + 1);
        int32_t vox[2] = {(int32_t)(cyl_pnt[0] / v_size), (int32_t)(cyl_pnt[1] / v_size)};
        int32_t x_min, x_max, y_min, y_max;
        if (calc) {
            x_min = imax(0, vox[0] - rad_vox * 20); x_max = imin((int32_t)res1, vox[0] + rad_vox * 20);
            y_min = imax(0, vox[1] - rad_vox * 20); y_max = imin((int32_t)res1, vox[1] + rad_vox * 20);
        } else {
            x_min = imax(0, vox[0] - rad_vox); x_max = imin((int32_t)res1, vox[0] + rad_vox + 2);
            y_min = imax(0, vox[1] - rad_vox); y_max = imin((int32_t)res1, vox[1] + rad_vox + 2);
        }
        for (int32_t pz = zlo; pz < zhi; pz++)
            for (int32_t py = y_min; py < y_max; py++)
                for (int32_t px = x_min; px < x_max; px++) {
                    size_t p = ((size_t)px * res1 + py) * nzw + (pz - zlo);
                    float p2p1[3] = {g[px] - cyl_pnt[0], g[py] - cyl_pnt[1], g[pz] - cyl_pnt[2]};
                    float distance2 = p2p1[0] * p2p1[0] + p2p1[1] * p2p1[1];
                    if (distance2 <= cyl_rad2) mask[p] = 1;
                    if (calc) {
                        /* cyl_dir = (0,0,1): dot = 0*p0 + 0*p1 + 1*p2; temp = dot*cyl_dir + cyl_pnt; perpendicular = grid - temp */
                        float dot = 0.0f * p2p1[0] + 0.0f * p2p1[1] + 1.0f * p2p1[2];
                        float temp[3] = {dot * 0.0f + cyl_pnt[0], dot * 0.0f + cyl_pnt[1], dot * 1.0f + cyl_pnt[2]};
                        float perp[3] = {g[px] - temp[0], g[py] - temp[1], g[pz] - temp[2]};
                        float nrm = sqrtf(perp[0] * perp[0] + perp[1] * perp[1] + perp[2] * perp[2]);
                        float phi_c = (perp[0] * B0_prj[0] + perp[1] * B0_prj[1] + perp[2] * B0_prj[2]) / nrm;
                        float phi_2c2_1 = 2 * phi_c * phi_c - 1;
                        if (distance2 > cyl_rad2)
                            fieldmap[p] = (float)(fieldmap[p] + 2 * M_PI * (1 - Y) * dChi * (cyl_rad2 / distance2) * phi_2c2_1 * theta_s2);
                        else
                            fieldmap[p] = (float)(fieldmap[p] + 2 * M_PI * (1 - Y) * dChi * (theta_c2 - 1.0 / 3.0));
                    }
                }
    }
}

/* ---- spheres ---- */
/* [phantom_sphere.cpp:22-55], serial order */
static int sph_overlap(const float *pts, const float *radii, size_t n, const float *sph_pnt, float *radius, int is_random_radius)
{
    for (size_t c = 0; c < n; c++) {
        float p0 = sph_pnt[0] - pts[3 * c], p1 = sph_pnt[1] - pts[3 * c + 1], p2 = sph_pnt[2] - pts[3 * c + 2];
        float distance = sqrtf(p0 * p0 + p1 * p1 + p2 * p2);
        if (distance <= radii[c] || distance <= *radius) return 1;
        else if (distance < radii[c] + *radius) {
            if (!is_random_radius) return 1;
            *radius = distance - radii[c];
        }
    }
    return 0;
}

/* [phantom_sphere.cpp:79-119] */
static int sph_place(const swo_phantom_spec *s, float **pts_out, float **radii_out, uint32_t *n_out)
{
    const float fov = s->fov_um, m_radius = s->radius_um, vf = s->volume_fraction;
    if (2 * m_radius >= fov) return 1;
    const int is_random_radius = m_radius < 0;
    const float max_radius = m_radius > 0 ? m_radius : -m_radius;
    size_t cap = 1024, n = 0;
    float *pts = malloc(cap * 3 * sizeof(float)), *radii = malloc(cap * sizeof(float));
    urng_t gen;
    gen.kind = 1;
    lcg_seed(&gen.lcg, (uint64_t)s->seed);
    float sph_pnt[3], radius, vol_sph = 0, vol_tol = fov * fov * fov;
    int32_t progress = 0;
    while (progress < 100) {
        radius = is_random_radius ? canonical(&gen) * max_radius : max_radius;
        for (int i = 0; i < 3; i++) sph_pnt[i] = canonical(&gen) * fov;
        if (sph_overlap(pts, radii, n, sph_pnt, &radius, is_random_radius)) continue;
        vol_sph = (float)(vol_sph + 4 * M_PI / 3 * radius * radius * radius);
        if (n == cap) {
            cap *= 2;
            pts = realloc(pts, cap * 3 * sizeof(float));
            radii = realloc(radii, cap * sizeof(float));
        }
        memcpy(pts + 3 * n, sph_pnt, 3 * sizeof(float));
        radii[n++] = radius;
        progress = (int32_t)(0.95 * 100 * (100. * vol_sph / vol_tol / vf));
    }
    *pts_out = pts;
    *radii_out = radii;
    *n_out = (uint32_t)n;
    return 0;
}

/* [phantom_sphere.cpp:121-198]; B0 = (0,0,1) (never rotated for spheres, phantom_base.h:50) */
static void sph_fill(const swo_phantom_spec *s, const float *g, const float *pts, const float *radii, uint32_t n, uint8_t *mask, float *fieldmap,
                     int32_t zlo, int32_t zhi)
{
    const size_t res1 = s->resolution, nzw = (size_t)(zhi - zlo);
    const int calc = fieldmap != NULL;
    const float fov = s->fov_um, Y = s->Y, dChi = s->dchi;
    const float B0[3] = {0.f, 0.f, 1.f};
    float v_size = fov / res1;
    for (uint32_t c = 0; c < n; c++) {
        const float *ctr = pts + 3 * c;
        float sph_rad = radii[c], sph_rad2 = sph_rad * sph_rad;
        int32_t rad_vox = (int32_t)(ceilf(sph_rad / v_size) + 1);
        int32_t lo[3], hi[3];
        for (int i = 0; i < 3; i++) {
            int32_t v = (int32_t)(ctr[i] / v_size);
            if (calc) { lo[i] = imax(0, v - rad_vox * 20); hi[i] = imin((int32_t)res1, v + rad_vox * 20); }
            else { lo[i] = imax(0, v - rad_vox); hi[i] = imin((int32_t)res1, v + rad_vox + 2); }
        }
        for (int32_t pz = imax(lo[2], zlo); pz < imin(hi[2], zhi); pz++)
            for (int32_t py = lo[1]; py < hi[1]; py++)
                for (int32_t px = lo[0]; px < hi[0]; px++) {
                    size_t p = ((size_t)px * res1 + py) * nzw + (pz - zlo);
                    float p2p1[3] = {g[px] - ctr[0], g[py] - ctr[1], g[pz] - ctr[2]};
                    float distance2 = p2p1[0] * p2p1[0] + p2p1[1] * p2p1[1] + p2p1[2] * p2p1[2];
                    if (distance2 <= sph_rad2) mask[p] = 1;
                    if (calc) {
                        float dp = p2p1[0] * B0[0] + p2p1[1] * B0[1] + p2p1[2] * B0[2];
                        float phi_c2 = dp * dp / distance2;
                        fieldmap[p] = (float)(fieldmap[p] + (distance2 > sph_rad2
                                          ? 4 * M_PI * (1 - Y) * dChi * sph_rad2 * sph_rad / distance2 / sqrtf(distance2) * (phi_c2 - 1. / 3.)
                                          : 0.f));
                    }
                }
    }
}

/* std::accumulate(m_mask.begin(), m_mask.end(), 0) * 100.0 / m_mask.size() -> float  [phantom_cylinder.cpp:270] */
static float volume_fraction(const uint8_t *mask, size_t V)
{
    int acc = 0;
    for (size_t i = 0; i < V; i++) acc += (int8_t)mask[i];
    return (float)(acc * 100.0 / V);
}

int swo_phantom_shapes(const swo_phantom_spec *s, float *shapes, uint32_t cap, uint32_t *n_shapes)
{
    float *pts = NULL, *radii = NULL;
    uint32_t n = 0;
    int rc = 0;
    if (s->shape == SWO_SHAPE_CYLINDER) {
        float *g = malloc(s->resolution * sizeof(float));
        grid_base(s->fov_um, s->resolution, g);
        rc = cyl_place(s, g, &pts, &radii, &n);
        free(g);
    } else if (s->shape == SWO_SHAPE_SPHERE) rc = sph_place(s, &pts, &radii, &n);
    if (rc) return rc;
    for (uint32_t i = 0; shapes && i < n && i < cap; i++) {
        memcpy(shapes + 4 * i, pts + 3 * i, 3 * sizeof(float));
        shapes[4 * i + 3] = radii[i];
    }
    *n_shapes = n;
    free(pts);
    free(radii);
    return 0;
}

int swo_phantom_generate_window(const swo_phantom_spec *s, int32_t zlo, int32_t zhi, uint8_t *mask, float *fieldmap, float *shapes, uint32_t cap,
                                uint32_t *n_shapes)
{
    const size_t res = s->resolution;
    const int calc = s->Y >= 0 && s->shape != SWO_SHAPE_TWOPOOLS;
    if (s->fov_um == 0 || res == 0) return 1; /* phantom_base.cpp:110-114 */
    if (calc && !fieldmap) return 2;
    if (zlo < 0 || zhi > (int32_t)res || zlo >= zhi) return 3;
    const size_t W = res * res * (size_t)(zhi - zlo);
    memset(mask, 0, W);
    if (calc) memset(fieldmap, 0, W * sizeof(float));
    if (n_shapes) *n_shapes = 0;
    if (s->shape == SWO_SHAPE_TWOPOOLS) { /* phantom_twopools.cpp:55: first half of the flat [x][y][z] array = 1 */
        const size_t V = res * res * res;
        for (size_t x = 0; x < res; x++)
            for (size_t y = 0; y < res; y++)
                for (int32_t z = zlo; z < zhi; z++)
                    mask[(x * res + y) * (size_t)(zhi - zlo) + (size_t)(z - zlo)] = ((x * res + y) * res + (size_t)z) < V / 2;
        return 0;
    }
    float *g = malloc(res * sizeof(float));
    grid_base(s->fov_um, res, g);
    float *pts = NULL, *radii = NULL;
    uint32_t n = 0;
    int rc = s->shape == SWO_SHAPE_CYLINDER ? cyl_place(s, g, &pts, &radii, &n) : sph_place(s, &pts, &radii, &n);
    if (rc) { free(g); return rc; }
    if (s->shape == SWO_SHAPE_CYLINDER) cyl_fill(s, g, pts, radii, n, mask, calc ? fieldmap : NULL, zlo, zhi);
    else sph_fill(s, g, pts, radii, n, mask, calc ? fieldmap : NULL, zlo, zhi);
    for (uint32_t i = 0; shapes && i < n && i < cap; i++) {
        memcpy(shapes + 4 * i, pts + 3 * i, 3 * sizeof(float));
        shapes[4 * i + 3] = radii[i];
    }
    if (n_shapes) *n_shapes = n;
    free(g);
    free(pts);
    free(radii);
    return 0;
}

int swo_phantom_generate(const swo_phantom_spec *s, uint8_t *mask, float *fieldmap, float *bvf, float *shapes, uint32_t cap, uint32_t *n_shapes)
{
    const size_t res = s->resolution;
    int rc = swo_phantom_generate_window(s, 0, (int32_t)res, mask, fieldmap, shapes, cap, n_shapes);
    if (rc == 0 && bvf) *bvf = volume_fraction(mask, res * res * res);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------------------------
 * Triangle-mesh phantom (`spinwalk phantom -p -i mesh.ply`): src/phantom/phantom_ply.cpp:141-227 + helpers :27-134.
 * A voxel centre is inside the mesh when the ray (1,0,0) from it hits an odd number of triangles (Möller-Trumbore in mixed
 * double / float arithmetic, :90-111).  The reference walks a median-split BVH whose boxes only prune on (y, z) (:37-88,
 * phantom_ply.h:37-42), so a triangle is tested for a ray exactly when the ray's (y, z) lies inside the box of the LEAF that
 * holds the triangle (a leaf box lies inside all its ancestors' boxes).  The BVH is rebuilt here with the same median splits
 * (qsort instead of std::sort: triangles with equal centroid keys may land in different leaves, which can only matter for a ray
 * through the exact edge of a leaf box); the hit count then runs over all triangles with the leaf-box test in place.
 * vertices: double [nv][3] as the PLY file holds them (mm); faces: [nf][3] vertex indices.
 * ------------------------------------------------------------------------------------------------------------------------ */
typedef struct { double x, y, z; } vec3;
typedef struct { vec3 v0, v1, v2; double ymin, ymax, zmin, zmax; } mtri;

static int g_axis;
static double coord(const vec3 *v, int ax) { return ax == 0 ? v->x : ax == 1 ? v->y : v->z; }
static int cmp_centroid(const void *pa, const void *pb)
{ /* :62-66: (c0 + c1 + c2) / 3.0f in double */
    const mtri *a = pa, *b = pb;
    double ca = (coord(&a->v0, g_axis) + coord(&a->v1, g_axis) + coord(&a->v2, g_axis)) / 3.0f;
    double cb = (coord(&b->v0, g_axis) + coord(&b->v1, g_axis) + coord(&b->v2, g_axis)) / 3.0f;
    return ca < cb ? -1 : (cb < ca ? 1 : 0);
}
static void tri_box(const mtri *t, size_t start, size_t end, vec3 *mn, vec3 *mx)
{ /* computeAABB :43-52 */
    mn->x = mn->y = mn->z = 1.7976931348623157e308;
    mx->x = mx->y = mx->z = -1.7976931348623157e308;
    for (size_t i = start; i < end; i++) {
        const vec3 *v[3] = {&t[i].v0, &t[i].v1, &t[i].v2};
        for (int k = 0; k < 3; k++) {
            if (v[k]->x < mn->x) mn->x = v[k]->x;
            if (v[k]->y < mn->y) mn->y = v[k]->y;
            if (v[k]->z < mn->z) mn->z = v[k]->z;
            if (v[k]->x > mx->x) mx->x = v[k]->x;
            if (v[k]->y > mx->y) mx->y = v[k]->y;
            if (v[k]->z > mx->z) mx->z = v[k]->z;
        }
    }
}
static void build_leaves(mtri *t, size_t start, size_t end)
{ /* buildBVH :73-86 + partitionTriangles :54-71 */
    vec3 mn, mx;
    tri_box(t, start, end, &mn, &mx);
    if (end - start <= 4) {
        for (size_t i = start; i < end; i++) { t[i].ymin = mn.y; t[i].ymax = mx.y; t[i].zmin = mn.z; t[i].zmax = mx.z; }
        return;
    }
    double ex = mx.x - mn.x, ey = mx.y - mn.y, ez = mx.z - mn.z;
    int axis = 0;
    if (ey > ex) axis = 1;
    if (ez > (ex > ey ? ex : ey)) axis = 2;
    g_axis = axis;
    qsort(t + start, end - start, sizeof(mtri), cmp_centroid);
    size_t mid = start + (end - start) / 2;
    build_leaves(t, start, mid);
    build_leaves(t, mid, end);
}

/* rayIntersectsTriangle :90-111 with dir = (1,0,0); every product / sum in the reference's type and order */
static int ray_hits(const vec3 *o, const mtri *tr)
{
    const float EPSILON = 1e-6f;
    const vec3 dir = {1., 0., 0.};
    vec3 e1 = {tr->v1.x - tr->v0.x, tr->v1.y - tr->v0.y, tr->v1.z - tr->v0.z};
    vec3 e2 = {tr->v2.x - tr->v0.x, tr->v2.y - tr->v0.y, tr->v2.z - tr->v0.z};
    vec3 h = {dir.y * e2.z - dir.z * e2.y, dir.z * e2.x - dir.x * e2.z, dir.x * e2.y - dir.y * e2.x};
    float a = (float)(e1.x * h.x + e1.y * h.y + e1.z * h.z);
    if (fabsf(a) < EPSILON) return 0;
    float f = 1.0f / a;
    vec3 s = {o->x - tr->v0.x, o->y - tr->v0.y, o->z - tr->v0.z};
    float u = f * (float)(s.x * h.x + s.y * h.y + s.z * h.z);
    if (u < 0.0f || u > 1.0f) return 0;
    vec3 q = {s.y * e1.z - s.z * e1.y, s.z * e1.x - s.x * e1.z, s.x * e1.y - s.y * e1.x};
    float v = f * (float)(dir.x * q.x + dir.y * q.y + dir.z * q.z);
    if (v < 0.0f || u + v > 1.0f) return 0;
    float t = f * (float)(e2.x * q.x + e2.y * q.y + e2.z * q.z);
    return t >= 0.0f;
}

int swo_phantom_mesh(float fov_um, uint64_t resolution, const double *vertices, uint64_t n_vertices, const uint64_t *faces, uint64_t n_faces, uint8_t *mask)
{
    const size_t res = resolution;
    if (fov_um == 0 || res == 0) return 1;
    memset(mask, 0, res * res * res);
    mtri *t = malloc((n_faces ? n_faces : 1) * sizeof(mtri));
    for (uint64_t i = 0; i < n_faces; i++) { /* :160-167: mm -> um */
        const uint64_t *f = faces + 3 * i;
        if (f[0] >= n_vertices || f[1] >= n_vertices || f[2] >= n_vertices) { free(t); return 2; }
        t[i].v0 = (vec3){vertices[3 * f[0]] * 1e3, vertices[3 * f[0] + 1] * 1e3, vertices[3 * f[0] + 2] * 1e3};
        t[i].v1 = (vec3){vertices[3 * f[1]] * 1e3, vertices[3 * f[1] + 1] * 1e3, vertices[3 * f[1] + 2] * 1e3};
        t[i].v2 = (vec3){vertices[3 * f[2]] * 1e3, vertices[3 * f[2] + 1] * 1e3, vertices[3 * f[2] + 2] * 1e3};
    }
    vec3 mn, mx;
    tri_box(t, 0, n_faces, &mn, &mx);
    vec3 shift = {(mx.x + mn.x) / 2., (mx.y + mn.y) / 2., (mx.z + mn.z) / 2.}; /* :172: centre the mesh in the FoV */
    const double half = fov_um / 2.;
    for (uint64_t i = 0; i < n_faces; i++) {
        vec3 *v[3] = {&t[i].v0, &t[i].v1, &t[i].v2};
        for (int k = 0; k < 3; k++) { v[k]->x = v[k]->x + half - shift.x; v[k]->y = v[k]->y + half - shift.y; v[k]->z = v[k]->z + half - shift.z; }
    }
    build_leaves(t, 0, n_faces);
    tri_box(t, 0, n_faces, &mn, &mx); /* root->bounds */
    float *g = malloc(res * sizeof(float));
    grid_base(fov_um, res, g);
    for (size_t px = 0; px < res; px++)
        for (size_t py = 0; py < res; py++)
            for (size_t pz = 0; pz < res; pz++) {
                vec3 o = {g[px], g[py], g[pz]};
                if (o.x < mn.x || o.x > mx.x || o.y < mn.y || o.y > mx.y || o.z < mn.z || o.z > mx.z) continue; /* :206-207 */
                unsigned hits = 0;
                for (uint64_t i = 0; i < n_faces; i++) {
                    if (o.y < t[i].ymin || o.y > t[i].ymax || o.z < t[i].zmin || o.z > t[i].zmax) continue; /* leaf box (intersectsRay) */
                    hits += (unsigned)ray_hits(&o, &t[i]);
                }
                mask[(px * res + py) * res + pz] = (uint8_t)(hits % 2);
            }
    free(g);
    free(t);
    return 0;
}
