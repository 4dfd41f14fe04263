/* include/spinwalk_engine.h — C-ABI of the B200-native SpinWalk `sim` engine (libspinwalk_b200.so).
 *
 * This is the drop-in boundary for ONE path of aghaeifar/SpinWalk v1.21.0: the per-spin Monte-Carlo
 * time loop.  The reference has no plugin/FFI interface for it; the seam it replaces is the device
 * branch of sim::monte_carlo::run,
 *      src/sim/monte_carlo.cu:247-262   device allocation + thrust H2D uploads
 *      src/sim/monte_carlo.cu:273-337   per-scale loop, cu_sim<<<>>> launch + sync
 *      src/sim/monte_carlo.cu:170-176   thrust D2H copies in save()
 * i.e. everything between "host vectors are filled" and "host vectors hold the results".
 * INTEGRATION.md shows the patch a reference maintainer would apply.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no torch / thrust / STL types cross this boundary;
 *   - every function returns an int status (SWK_OK == 0); swk_last_error() gives the message,
 *     mirroring the reference's `bool` + log-line convention (monte_carlo.cu:203-206,229-233);
 *   - the caller owns all host buffers, the engine owns all device memory;
 *   - an engine handle drives one CUDA device and is not thread-safe (the reference drives one
 *     device from one host thread too, monte_carlo.cu:44-52);
 *   - there is NO CPU fallback: without a usable CUDA device swk_create fails.
 */
#ifndef SPINWALK_ENGINE_H
#define SPINWALK_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWK_VERSION_MAJOR 0
#define SWK_VERSION_MINOR 1

enum swk_status {
    SWK_OK = 0,
    SWK_ERR_INVALID = 1,   /* bad argument / inconsistent tables (≙ config_reader::check, config_reader.cpp:195-324) */
    SWK_ERR_CUDA = 2,      /* CUDA runtime error                                                                     */
    SWK_ERR_MEMORY = 3,    /* would not fit in free device memory (≙ check_memory_size, device_helper.cu:78-101)     */
    SWK_ERR_STATE = 4,     /* call order: phantom / sequence / spins not set                                         */
    SWK_ERR_SUBSTRATE = 5  /* mask holds more substrates than the sequence (≙ monte_carlo.cu:113-118)                */
};

/* WHAT_TO_SCALE (config_reader.h:12, monte_carlo.cu:277-305) */
enum swk_scale_type { SWK_SCALE_FOV = 0, SWK_SCALE_GRADIENT = 1, SWK_SCALE_PHASE_CYCLING = 2 };

/* Arithmetic of the walk.
 *   SWK_MODE_COMPAT  the reference CUDA build's arithmetic: thrust::minstd_rand seeded seed+spin and
 *                    discarded seed+spin (kernels.cu:77-88), normal = -sqrt(2) erfcinvf(2p), FP64 positions
 *                    in metres.  Bit-exact walks against the reference's own cu_sim on the same device.
 *   SWK_MODE_FAST    Philox4x32-10 counter RNG (fixed key; counter = Philox block index, seed, global spin id, stream tag):
 *                    round r of a spin draws its three normals from block r >> 1 (Box-Muller on 23-bit uniforms), and every
 *                    scale replays the same stream, like the reference re-seeding seed+spin per scale (kernels.cu:77-88); the
 *                    permeability uniform is a separate Philox2x32-10 stream keyed by a fold of the seed, counter = (round, spin id).
 *                    Positions are 32-bit fixed-point grid coordinates.  Same stochastic process; agrees with the reference
 *                    within Monte-Carlo error (tests/test_fast_parity_gpu.py).  Deviations from the reference's arithmetic:
 *                    DESIGN.md §2. */
enum swk_mode { SWK_MODE_COMPAT = 0, SWK_MODE_FAST = 1 };

/* ≙ struct parameters AFTER parameters::prepare (simulation_parameters.cuh:176-201,227-245).
 * fov / phantom_size / matrix_length / fieldmap_exist of the reference POD are set by swk_set_phantom. */
typedef struct swk_params {
    float    B0;                      /* T;  fieldmap (Tesla at 1 T) is scaled by B0*dt*gamma*180/pi in the gather */
    float    c, s;                    /* cosf / sinf of RF_FA[0]                (simulation_parameters.cuh:229-230)  */
    float    linear_phase_cycling;    /* deg                                                                        */
    float    quadratic_phase_cycling; /* deg                                                                        */
    int32_t  timestep_us;
    int32_t  TR_us;
    int32_t  n_dummy_scan;            /* >= 0 (the "-1 => 5 T1/TR" rule is parameters::prepare's, see swk_prepare)  */
    uint32_t n_spins;                 /* GLOBAL spin count: denominator of the DEPHASING term (kernels.cu:176)      */
    uint32_t n_timepoints;            /* TR_us / timestep_us                                                        */
    uint32_t n_substrate;
    uint64_t seed;                    /* non-zero                                                                   */
    uint64_t max_iterations;
    int32_t  cross_fov;               /* CROSS_FOV                                                                  */
    int32_t  record_trajectory;       /* RECORD_TRAJECTORY                                                          */
} swk_params;

/* ≙ struct parameters_uvec (simulation_parameters.cuh:94-173): 14 non-owning {ptr,len} views of HOST
 * arrays.  *_tp are in TIMEPOINTS (config_reader.cpp:39-46); step_sigma_m is parameters_hvec::diffusivity
 * after prepare, i.e. 1e-3*sqrt(2*D*dt_us) metres (simulation_parameters.cuh:236-237). */
typedef struct swk_tables {
    const double  *step_sigma_m;  uint32_t n_step_sigma;   /* [n_substrate]               */
    const float   *T1_ms;         uint32_t n_T1;           /* [n_substrate]               */
    const float   *T2_ms;         uint32_t n_T2;           /* [n_substrate]               */
    const float   *pXY;           uint32_t n_pXY;          /* [n_substrate^2] [from][to]  */
    const float   *RF_FA_deg;     uint32_t n_RF_FA;
    const float   *RF_PH_deg;     uint32_t n_RF_PH;
    const int32_t *RF_tp;         uint32_t n_RF;           /* RF_tp[0] == 0               */
    const int32_t *TE_tp;         uint32_t n_TE;
    const float   *dephasing_deg; uint32_t n_dephasing_deg;
    const int32_t *dephasing_tp;  uint32_t n_dephasing;
    const float   *gradX_mTm;     uint32_t n_gradX;
    const float   *gradY_mTm;     uint32_t n_gradY;
    const float   *gradZ_mTm;     uint32_t n_gradZ;
    const int32_t *gradient_tp;   uint32_t n_gradient;
} swk_tables;

/* Work counters of the last run, summed over scales and this engine's spins. */
typedef struct swk_stats {
    uint64_t steps;          /* accepted spin-steps                                          */
    uint64_t mask_gathers;   /* iterations whose voxel index changed  (kernels.cu:150)       */
    uint64_t field_gathers;  /* accepted voxel changes                (kernels.cu:165)       */
    uint64_t rejects;        /* permeability rejections               (kernels.cu:154)       */
    uint64_t lost;           /* spins abandoned: out of range or stuck (kernels.cu:141-159)  */
    float    kernel_ms;      /* device time of the walk kernel(s), CUDA events on the engine stream */
    float    device_ms;      /* device time of the whole pass: output zero-fill + kernels           */
    uint32_t n_launches;     /* kernels of this library launched by the run                  */
} swk_stats;

typedef struct swk_engine swk_engine;

/* ---- lifetime (≙ monte_carlo ctor: check_CUDA, device count, cudaSetDevice; monte_carlo.cu:33-55) ---- */
int         swk_create(int device_id, swk_engine **out);
void        swk_destroy(swk_engine *e);
const char *swk_last_error(const swk_engine *e); /* e may be NULL: error of the last failed swk_create */
int         swk_version(void);                   /* major*100 + minor */
int         swk_device_count(void);              /* ≙ sim::get_device_count, device_helper.cu; <=0 if none */
/* ≙ sim::print_device_info (device_helper.cu:47-75, the `-g` flag): the same lines — driver / runtime CUDA versions, device count,
 * name, compute capability, free and total memory of the current device — written into buf (NUL-terminated, truncated to n). */
int         swk_device_info(char *buf, size_t n);

/* ---- parameters::prepare (simulation_parameters.cuh:227-245) for callers that hold raw INI values ----
 * Fills p->c, p->s, p->n_timepoints, resolves n_dummy_scan < 0, converts diffusivity_m2s[n] -> sigma_out[n]. */
int swk_prepare(swk_params *p, float RF_FA0_deg, float T1_0_ms, const double *diffusivity_m2s, uint32_t n, double *sigma_out);

/* ---- phantom (≙ read_phantom outputs + d_mask/d_fieldmap upload; monte_carlo.cu:98-123,254,257) ----
 * mask: uint8 [dims0][dims1][dims2] row-major, x slowest (kernels.cuh:53-60).  fieldmap_T: float, same
 * shape, Tesla at B0 = 1 T, or NULL.  fov_m: metres.  Pointers are HOST pointers unless on_device != 0
 * (then they are device pointers on this engine's device and are copied device-to-device).
 * Frees the previous phantom first (≙ cleanup_device, monte_carlo.cu:86-95). */
int swk_set_phantom(swk_engine *e, const uint8_t *mask, const float *fieldmap_T, const uint64_t dims[3],
                    const float fov_m[3], int on_device);

/* ---- sequence + tissue tables (≙ param_dvec.copy_from_host + kernel arguments; monte_carlo.cu:219-225) ---- */
int swk_set_sequence(swk_engine *e, const swk_params *p, const swk_tables *t);

/* ---- spins of this engine's shard: global ids [spin_first, spin_first + n_local) ----
 * XYZ0: host float [n_local][3], metres, UNSCALED (FoV scaling happens on the device, ≙ monte_carlo.cu:278).
 *       NULL => positions are generated on the device, uniform in [1%,99%] of the FoV
 *       (same distribution as monte_carlo.cu:142-151, Philox stream keyed by seed and global id).
 * M0:   host float [n_local][3] or NULL => (0,0,1) (monte_carlo.cu:162-164). */
int swk_set_spins(swk_engine *e, const float *XYZ0, const float *M0, uint32_t spin_first, uint32_t n_local);

/* ---- run all scales with inputs resident on the device (≙ the scale loop, monte_carlo.cu:273-337) ----
 * scales: host float [n_scales].  flags: SWK_OUT_*.  d_sums: DEVICE pointer to double
 * [n_scales][n_TE][n_substrate][4] = {sum Mx, sum My, sum Mz, count} per tissue at each echo, or NULL
 * (engine-owned buffer is used; read it with swk_get_sums).  The buffer is overwritten by the run.  The sums are accumulated in
 * integer fixed point (2^-22 per component and spin): they are bit-reproducible run to run and independent of how the spins are
 * sharded, sliced or ordered. */
enum { SWK_OUT_M1 = 1, SWK_OUT_XYZ1 = 2, SWK_OUT_T = 4, SWK_OUT_ALL = 7,
       SWK_RUN_STATS = 16,  /* count gathers / rejections (swk_stats); slightly slower kernel variant */
       SWK_RUN_NO_SORT = 32, /* simulate spins in caller order (no substrate/Morton locality order); for A/B tests */
       SWK_RUN_NO_PACK = 64, /* FAST mode: gather mask byte + FP32 field separately instead of the packed voxel word */
       SWK_RUN_NO_REBIN = 128, /* FAST mode: never pause a long run (many TRs) to re-sort the spins by their current voxel; for A/B tests */
       SWK_RUN_ZSLAB = 256, /* accepted and ignored (round 1's opt-in; the z-slab table is the default now, see SWK_RUN_NO_ZSLAB) */
       SWK_RUN_NO_ZSLAB = 512, /* FAST mode: by default a phantom whose mask and field map do not depend on z (checked once per phantom on the
                              device; every cylinder phantom of `spinwalk phantom -c`) is walked on the packed voxel words of ONE z plane — the
                              same words, hence the same results bit for bit, from an L1/L2-resident [nx][ny] table instead of [nx][ny][nz].
                              This flag (or environment SWK_NO_ZSLAB=1) keeps the full table: A/B tests, the gather-roofline measurement. */
       SWK_RUN_NO_SHARE = 1024, /* FAST mode: every thread generates its own normals (the PRIVATE kernel variant) even when several scales could
                              share one generation per spin (walk_fast.cuh); same results bit for bit; for A/B tests (also: SWK_NO_SHARE=1) */
       SWK_RUN_NO_ONEWALK = 2048 /* FAST mode: when the scales act on the gradients or on the phase cycling (WHAT_TO_SCALE 1, 2) every scale of a spin
                              walks the same path (the reference re-seeds seed+spin per scale, kernels.cu:77-88, and the FoV is not scaled), so by
                              default ONE walker per spin carries the magnetisation of every scale: the walk is paid once instead of n_scales
                              times, positions and tissues are bit-identical, magnetisations agree to FP32 round-off (the gradient phase is
                              scaled after the sum instead of term by term).  swk_stats then counts every step / gather / rejection n_scales
                              times (the statistics of the walks it stands for).  Not used while trajectories are recorded.  This flag (or
                              SWK_NO_ONEWALK=1) walks every scale separately: A/B tests. */ };
int swk_run_device(swk_engine *e, const float *scales, uint32_t n_scales, int scale_type, int mode, int flags,
                   double *d_sums);

/* ---- results to the host (≙ thrust::copy D2H in save(), monte_carlo.cu:170-176) ----
 * Layouts are the reference's (monte_carlo.cu:61-70), restricted to this engine's spins:
 *   M1 float [n_scales][n_local][n_TE][3], XYZ1 float [n_scales][n_local][trj][3], T uint8 [n_scales][n_local][n_TE].
 * Any pointer may be NULL.  Unwritten entries (lost spins) read 0, as in the reference. */
int swk_download(swk_engine *e, float *M1, float *XYZ1, uint8_t *T);
int swk_get_sums(swk_engine *e, double *sums /* host [n_scales][n_TE][n_substrate][4] */);
int swk_get_stats(swk_engine *e, swk_stats *out);

/* ---- one-call convenience with HOST buffers: set_spins + run_device + download (+ sums) ---- */
int swk_run(swk_engine *e, const float *XYZ0, const float *M0, uint32_t spin_first, uint32_t n_local,
            const float *scales, uint32_t n_scales, int scale_type, int mode,
            float *M1, float *XYZ1, uint8_t *T, double *sums, swk_stats *stats);

/* ---- several engines filling ONE set of host arrays (one engine per GPU, spins sharded) ----
 * By default swk_download / swk_run treat the host arrays as this engine's own [n_scales][n_local][...] block.  After
 * swk_set_host_rows(e, n_rows_total, row_first) they are the GLOBAL arrays [n_scales][n_rows_total][...] of the reference
 * (monte_carlo.cu:61-70) and this engine writes rows [row_first, row_first + n_local) of every scale — each GPU copies its
 * contiguous slices straight into place, no exchange between GPUs.  n_rows_total = 0 restores the default. */
int swk_set_host_rows(swk_engine *e, uint64_t n_rows_total, uint64_t row_first);

/* ---- page-locked host memory for the caller's big arrays (optional; pageable buffers work, just slower) ---- */
int  swk_alloc_pinned(void **ptr, size_t bytes);
void swk_free_pinned(void *ptr);

/* ---- diagnostics: the roofline of the voxel fetch, measured on THIS device and THIS phantom ----
 * Launches a kernel that does nothing but dependent 4-byte gathers at uniformly random addresses of the engine's voxel
 * table (the one the last run walked: the z slab, else the packed words when they exist, else the fieldmap, else the mask read as words) with the walk's load
 * instruction, `threads_per_sm` resident threads per SM and `iters` gathers per thread, and reports gathers/s
 * (CUDA events on the engine stream).  What the walk can reach at small FoV scales, where every step lands in a
 * voxel far from the last one.  Not part of the reference (it has no profiling hooks). */
int swk_probe_gather(swk_engine *e, uint32_t threads_per_sm, uint32_t iters, double *gathers_per_s, uint64_t *table_bytes);

/* ---- plumbing for callers that share the device with the engine (PyTorch, NCCL) ---- */
/* Test hook: runs the random-number building blocks of SWK_MODE_FAST on caller-supplied inputs in[n][4], results in out[n][8].
 *   which = 0: the Philox4x32-10 block of the displacement stream (fixed key 243F6A88 85A308D3), counter = in[i][0..3] -> 4 words;
 *   which = 1: the Philox2x32-10 word of the permeability stream, counter = in[i][0..1], key = in[i][2] -> word, FP32 bits of its uniform [0,1);
 *   which = 2: the six Box-Muller normals made from the 128-bit block in[i][0..3] -> 6 FP32 bit patterns (step order x y z, x y z).
 * No counterpart in the reference (its generators are thrust::minstd_rand / std::mt19937, src/sim/kernels.cu:77-88). */
int swk_debug_rng(swk_engine *e, int which, const uint32_t *in, uint32_t n, uint32_t *out);

void    *swk_stream(swk_engine *e);            /* cudaStream_t the engine launches on                       */
double  *swk_device_sums(swk_engine *e);       /* engine-owned device sums buffer of the last run, or NULL  */
uint64_t swk_device_bytes(const swk_engine *e);/* device memory currently held (≙ get_total_memory)         */

#ifdef __cplusplus
}
#endif
#endif /* SPINWALK_ENGINE_H */
