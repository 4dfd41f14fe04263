/* oracle/sim_case.h — TEST INFRASTRUCTURE, not product code.
 *
 * One plain-C description of a `sim` case, shared by
 *   - oracle/sim_oracle.c      (our CPU restatement of the reference hot path), and
 *   - oracle/ref_harness.cpp   (a caller of the UNMODIFIED reference sim::sim / cu_sim,
 *                               compiled into oracle/_ref/ from /root/reference).
 * Both take the same struct so tests can feed identical inputs to both.
 *
 * Field meanings follow the reference after config_reader::prepare but BEFORE
 * parameters::prepare (src/sim/simulation_parameters.cuh:227-245):
 *   - *_tp tables are in TIMEPOINTS (config_reader.cpp:39-46 already divided by TIME_STEP)
 *   - diffusivity is in m^2/s (parameters::prepare converts it to the per-axis step sigma)
 *   - n_dummy_scan < 0 means "5*T1[0]/TR" (simulation_parameters.cuh:239-242)
 *   - fieldmap passed next to this struct is in Tesla at B0 = 1 T (monte_carlo.cu:241-244
 *     converts to degrees per timestep; both implementations do that internally on a copy)
 */
#ifndef SWO_SIM_CASE_H
#define SWO_SIM_CASE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SWO_SCALE_FOV = 0, SWO_SCALE_GRADIENT = 1, SWO_SCALE_PHASE_CYCLING = 2 };

/* RNG flavours of the reference (src/sim/kernels.cu:76-88) */
enum {
    SWO_RNG_MT19937 = 0, /* host build: std::mt19937 + libstdc++ polar normal_distribution<float>   */
    SWO_RNG_MINSTD  = 1  /* CUDA build: thrust::minstd_rand + -sqrt(2)*erfcinv(2p) normal           */
};

typedef struct swo_case {
    double   fov[3];          /* metres, unscaled (monte_carlo.cu:264-265)                         */
    uint64_t phantom_size[3]; /* voxels, row-major x slowest (kernels.cuh:53-60)                   */
    uint64_t seed;            /* parameters::seed (must be non-zero; 0 means random_device)        */
    uint64_t max_iterations;
    float    B0;
    float    linear_phase_cycling, quadratic_phase_cycling;
    int32_t  timestep_us, TR_us, n_dummy_scan;
    uint32_t n_spins, n_substrate;
    int32_t  cross_fov, record_trajectory;

    const double  *diffusivity;          /* [n_substrate] m^2/s                                    */
    const float   *T1_ms, *T2_ms;        /* [n_substrate]                                          */
    const float   *pXY;                  /* [n_substrate^2] row-major [from][to]                   */
    const float   *RF_FA_deg, *RF_PH_deg;
    const int32_t *RF_tp;
    uint32_t       n_RF;
    const int32_t *TE_tp;
    uint32_t       n_TE;
    const float   *dephasing_deg;
    const int32_t *dephasing_tp;
    uint32_t       n_dephasing;
    const float   *gradX_mTm, *gradY_mTm, *gradZ_mTm;
    const int32_t *gradient_tp;
    uint32_t       n_gradient;

    const float   *scales;               /* config_reader.h:40 keeps them as float                 */
    uint32_t       n_scales;
    int32_t        scale_type;           /* SWO_SCALE_*                                            */
} swo_case;

/* Output sizes (monte_carlo.cu:61-70):
 *   trj = record_trajectory ? n_timepoints*(n_dummy_scan+1) : 1
 *   M1   float  [n_scales][n_spins][n_TE][3]
 *   XYZ1 float  [n_scales][n_spins][trj][3]
 *   T    uint8  [n_scales][n_spins][n_TE]
 * All three must be zero-initialised by the caller (monte_carlo.cu:256,259-260). */

#ifdef __cplusplus
}
#endif
#endif
