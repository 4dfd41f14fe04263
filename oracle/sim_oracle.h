/* oracle/sim_oracle.h — TEST INFRASTRUCTURE, not product code.
 * C interface of the CPU restatement (oracle/sim_oracle.c) of SpinWalk's `sim` hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library. */
#ifndef SWO_SIM_ORACLE_H
#define SWO_SIM_ORACLE_H

#include "sim_case.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Work counters, summed over all scales and the simulated spins (used for the roofline's
 * algorithmic-bytes accounting). */
typedef struct swo_stats {
    uint64_t steps;         /* accepted steps (t advanced)                              */
    uint64_t mask_gathers;  /* iterations whose voxel index changed  (kernels.cu:150)   */
    uint64_t field_gathers; /* accepted voxel changes                (kernels.cu:165)   */
    uint64_t rejects;       /* permeability rejections               (kernels.cu:154)   */
    uint64_t lost;          /* spins that returned early             (kernels.cu:146,158)*/
} swo_stats;

/* Full run: all scales, spins [spin_begin, spin_end).  flavour = SWO_RNG_*.
 * M1 / XYZ1 / T must be zero-initialised, sizes as in sim_case.h.  stats / seconds nullable.
 * Returns 0 on success, <0 on invalid input. */
int swo_run(const swo_case *c, const float *fieldmap_T, const uint8_t *mask, const float *XYZ0, const float *M0,
            float *M1, float *XYZ1, uint8_t *T, uint32_t spin_begin, uint32_t spin_end, int flavour, int n_threads,
            swo_stats *stats, double *seconds);

/* monte_carlo.cu:142-151 default initial positions. */
void swo_init_positions(uint64_t seed, const float fov[3], uint32_t n_spins, float *XYZ0);

/* parameters::prepare outputs, for tests of the host logic. */
uint32_t swo_n_timepoints(const swo_case *c);
int32_t  swo_n_dummy_scan(const swo_case *c);
double   swo_step_sigma(double diffusivity, int32_t timestep_us);
float    swo_tesla_to_deg_per_step(float B0, int32_t timestep_us);

/* small pieces exposed for unit tests (mirror tests/test_kernel.cpp of the reference) */
int64_t swo_sub2ind(int64_t x, int64_t y, int64_t z, int64_t nx, int64_t ny, int64_t nz);
void    swo_xrot(float s, float c, const float *m0, float *m1);
void    swo_yrot(float s, float c, const float *m0, float *m1);
void    swo_zrot(float s, float c, const float *m0, float *m1);
void    swo_relax(float e1, float e2, const float *m0, float *m1);
void    swo_xrot_withphase(float s, float c, float phase_deg, const float *m0, float *m1);

/* RNG pieces exposed for unit tests */
void   swo_minstd_normals(uint64_t seed_plus_spin, uint32_t n, float *out);
void   swo_minstd_uniforms(uint64_t seed_plus_spin, uint32_t n, float *out);
void   swo_mt_normals(uint64_t seed_plus_spin, uint32_t n, float *out);
void   swo_mt_uniforms(uint64_t seed_plus_spin, uint32_t n, float *out);
double swo_erfcinv(double x);

#ifdef __cplusplus
}
#endif
#endif
