/* oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * A caller of the UNMODIFIED reference hot path.  It is compiled together with
 * /root/reference/src/sim/kernels.cu (where that file lies; nothing is copied) into
 *   oracle/_ref/libswref_cpu.so   g++  : reference's no-CUDA build (CMakeLists.txt:57-70)
 *                                        => std::mt19937 + libstdc++ normal_distribution
 *   oracle/_ref/libswref_cuda.so  nvcc : reference's CUDA build
 *                                        => thrust::minstd_rand + erfcinv normal; exports both the
 *                                        __host__ instantiation of sim::sim and a launcher for the
 *                                        reference's own __global__ cu_sim (needs a GPU).
 *
 * What this file restates (on raw arrays, because monte_carlo.cu itself needs HDF5 + Boost +
 * TBB which are absent here) is only the DRIVER around the kernel:
 *   - parameters::prepare call                          (monte_carlo.cu:205)
 *   - fieldmap Tesla -> degree/timestep                 (monte_carlo.cu:241-244)
 *   - the per-scale loop                                (monte_carlo.cu:264-337)
 *   - default XYZ0 / M0 initialisation                  (monte_carlo.cu:142-151,162-164)
 * The physics is the reference's own sim::sim / sim::cu_sim.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "sim/kernels.cuh" /* the reference's header (brings simulation_parameters.cuh) */
#include "definitions.h"   /* GAMMA */
#include "sim_case.h"

#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

namespace {

struct prepared {
    parameters      param;
    parameters_hvec hvec;
    parameters_uvec uvec;
    std::vector<float> gx0, gy0, gz0; /* unscaled gradients (monte_carlo.cu:215-217) */
    float lin0 = 0.f, quad0 = 0.f;
    size_t trj = 1;
};

template <class T>
std::vector<T> vec_of(const T *p, uint32_t n) { return p ? std::vector<T>(p, p + n) : std::vector<T>(); }

void fill(prepared &P, const swo_case &c, size_t matrix_length, bool fieldmap_exist)
{
    parameters &p = P.param;
    p.B0 = c.B0;
    p.linear_phase_cycling = c.linear_phase_cycling;
    p.quadratic_phase_cycling = c.quadratic_phase_cycling;
    p.timestep_us = c.timestep_us;
    p.TR_us = c.TR_us;
    p.n_dummy_scan = c.n_dummy_scan;
    p.n_spins = c.n_spins;
    p.n_substrate = c.n_substrate;
    p.n_scales = c.n_scales;
    p.seed = c.seed;
    p.max_iterations = c.max_iterations;
    p.enCrossFOV = c.cross_fov != 0;
    p.enRecordTrajectory = c.record_trajectory != 0;
    p.fieldmap_exist = fieldmap_exist;
    p.matrix_length = (int64_t)matrix_length;
    for (int i = 0; i < 3; i++) { p.phantom_size[i] = c.phantom_size[i]; p.fov[i] = (float)c.fov[i]; /* fov is read as float, monte_carlo.cuh:37 */ }

    parameters_hvec &h = P.hvec;
    h.diffusivity = vec_of(c.diffusivity, c.n_substrate);
    h.T1_ms = vec_of(c.T1_ms, c.n_substrate);
    h.T2_ms = vec_of(c.T2_ms, c.n_substrate);
    h.pXY = vec_of(c.pXY, c.n_substrate * c.n_substrate);
    h.RF_FA_deg = vec_of(c.RF_FA_deg, c.n_RF);
    h.RF_PH_deg = vec_of(c.RF_PH_deg, c.n_RF);
    h.RF_us = vec_of(c.RF_tp, c.n_RF);
    h.TE_us = vec_of(c.TE_tp, c.n_TE);
    h.dephasing_deg = vec_of(c.dephasing_deg, c.n_dephasing);
    h.dephasing_us = vec_of(c.dephasing_tp, c.n_dephasing);
    h.gradientX_mTm = vec_of(c.gradX_mTm, c.n_gradient);
    h.gradientY_mTm = vec_of(c.gradY_mTm, c.n_gradient);
    h.gradientZ_mTm = vec_of(c.gradZ_mTm, c.n_gradient);
    h.gradient_us = vec_of(c.gradient_tp, c.n_gradient);

    p.prepare(h); /* the reference's own unit conversions */
    P.uvec.copy_from_host(h);
    P.gx0 = h.gradientX_mTm; P.gy0 = h.gradientY_mTm; P.gz0 = h.gradientZ_mTm;
    P.lin0 = p.linear_phase_cycling; P.quad0 = p.quadratic_phase_cycling;
    P.trj = p.enRecordTrajectory ? (size_t)p.n_timepoints * (p.n_dummy_scan + 1) : 1;
}

/* monte_carlo.cu:241-244 */
std::vector<float> prescale_fieldmap(const parameters &p, const float *fieldmap_T, size_t n)
{
    std::vector<float> f;
    if (!fieldmap_T) return f;
    float k = p.B0 * p.timestep_us * 1e-6 * GAMMA * RAD2DEG;
    f.resize(n);
    for (size_t i = 0; i < n; i++) f[i] = fieldmap_T[i] * k;
    return f;
}

/* one entry of the scale loop, host-side state changes only (monte_carlo.cu:277-305) */
void apply_scale(prepared &P, const swo_case &c, float scale, const float *XYZ0, std::vector<float> &XYZ0_scaled)
{
    if (c.scale_type == SWO_SCALE_FOV) {
        for (size_t i = 0; i < XYZ0_scaled.size(); i++) XYZ0_scaled[i] = XYZ0[i] * scale;
        for (int i = 0; i < 3; i++) P.param.fov[i] = scale * (float)c.fov[i];
    } else if (c.scale_type == SWO_SCALE_GRADIENT) {
        for (size_t i = 0; i < P.gx0.size(); i++) {
            P.hvec.gradientX_mTm[i] = P.gx0[i] * scale;
            P.hvec.gradientY_mTm[i] = P.gy0[i] * scale;
            P.hvec.gradientZ_mTm[i] = P.gz0[i] * scale;
        }
    } else if (c.scale_type == SWO_SCALE_PHASE_CYCLING) {
        P.param.linear_phase_cycling = P.lin0 * scale;
        P.param.quadratic_phase_cycling = P.quad0;
    }
}

} // namespace

extern "C" {

/* 0 = mt19937 flavour (g++ build), 1 = minstd flavour (nvcc build) */
int swref_flavour(void)
{
#ifdef __CUDACC__
    return SWO_RNG_MINSTD;
#else
    return SWO_RNG_MT19937;
#endif
}

/* monte_carlo.cu:142-151 — std::mt19937(seed), three uniform_real_distribution<float> */
void swref_init_positions(uint64_t seed, const float fov[3], uint32_t n_spins, float *XYZ0)
{
    std::mt19937 gen(seed);
    std::uniform_real_distribution<float> dx(0.01 * fov[0], 0.99 * fov[0]);
    std::uniform_real_distribution<float> dy(0.01 * fov[1], 0.99 * fov[1]);
    std::uniform_real_distribution<float> dz(0.01 * fov[2], 0.99 * fov[2]);
    for (size_t i = 0; i < n_spins; i++) {
        XYZ0[3 * i + 0] = dx(gen);
        XYZ0[3 * i + 1] = dy(gen);
        XYZ0[3 * i + 2] = dz(gen);
    }
}

/* Host execution of the reference's sim::sim over all scales and spins [spin_begin, spin_end).
 * Returns 0; *seconds (nullable) = wall time of the scale loop, like monte_carlo.cu:271,340. */
int swref_run(const swo_case *c, const float *fieldmap_T, const uint8_t *mask, const float *XYZ0, const float *M0,
              float *M1, float *XYZ1, uint8_t *T, uint32_t spin_begin, uint32_t spin_end, int n_threads, double *seconds)
{
    size_t V = (size_t)c->phantom_size[0] * c->phantom_size[1] * c->phantom_size[2];
    prepared P;
    fill(P, *c, V, fieldmap_T != nullptr);
    std::vector<float> fm = prescale_fieldmap(P.param, fieldmap_T, V);
    std::vector<float> XYZ0_scaled(XYZ0, XYZ0 + 3 * (size_t)c->n_spins);
    const size_t nTE = c->n_TE, S = c->n_spins;
    if (n_threads < 1) n_threads = 1;

    auto t0 = std::chrono::high_resolution_clock::now();
    for (uint32_t k = 0; k < c->n_scales; k++) {
        apply_scale(P, *c, c->scales[k], XYZ0, XYZ0_scaled);
        float *m1 = M1 + 3 * nTE * S * k;
        float *x1 = XYZ1 + 3 * S * P.trj * k;
        uint8_t *t1 = T + nTE * S * k;
        std::atomic<uint32_t> next(spin_begin);
        auto work = [&]() {
            const uint32_t chunk = 64;
            for (;;) {
                uint32_t b = next.fetch_add(chunk);
                if (b >= spin_end) break;
                uint32_t e = std::min(spin_end, b + chunk);
                for (uint32_t s = b; s < e; s++)
                    sim::sim(P.param, P.uvec, fm.empty() ? nullptr : fm.data(), mask, M0, XYZ0_scaled.data(), m1, x1, t1, s);
            }
        };
        std::vector<std::thread> th;
        for (int i = 1; i < n_threads; i++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    return 0;
}

#ifdef __CUDACC__
/* Launch of the reference's own __global__ cu_sim exactly as monte_carlo.cu:324-333 does:
 * grid ceil(n/256) x block 256, one launch + device sync per scale.  *kernel_ms (nullable) = sum of
 * cudaEvent times around the launches.  Returns 0, or the cudaError_t value. */
#define SWREF_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
int swref_cuda_run(const swo_case *c, const float *fieldmap_T, const uint8_t *mask, const float *XYZ0, const float *M0,
                   float *M1, float *XYZ1, uint8_t *T, int device, float *kernel_ms)
{
    SWREF_CK(cudaSetDevice(device));
    size_t V = (size_t)c->phantom_size[0] * c->phantom_size[1] * c->phantom_size[2];
    prepared P;
    fill(P, *c, V, fieldmap_T != nullptr);
    std::vector<float> fm = prescale_fieldmap(P.param, fieldmap_T, V);
    const size_t nTE = c->n_TE, S = c->n_spins, K = c->n_scales;

    parameters_dvec dvec;
    dvec.copy_from_host(P.hvec);
    P.uvec.copy_from_device(dvec);
    thrust::device_vector<float> d_fm(fm.begin(), fm.end());
    thrust::device_vector<uint8_t> d_mask(mask, mask + V);
    thrust::device_vector<float> d_M0(M0, M0 + 3 * S), d_XYZ0(XYZ0, XYZ0 + 3 * S);
    thrust::device_vector<float> d_M1(3 * nTE * S * K, 0.f), d_XYZ1(3 * S * P.trj * K, 0.f);
    thrust::device_vector<uint8_t> d_T(nTE * S * K, 0);
    std::vector<float> XYZ0_scaled(XYZ0, XYZ0 + 3 * S);

    cudaEvent_t e0, e1;
    SWREF_CK(cudaEventCreate(&e0));
    SWREF_CK(cudaEventCreate(&e1));
    float total = 0.f;
    for (uint32_t k = 0; k < K; k++) {
        apply_scale(P, *c, c->scales[k], XYZ0, XYZ0_scaled);
        if (c->scale_type == SWO_SCALE_FOV) d_XYZ0 = XYZ0_scaled;
        if (c->scale_type == SWO_SCALE_GRADIENT) {
            dvec.gradientX_mTm = P.hvec.gradientX_mTm;
            dvec.gradientY_mTm = P.hvec.gradientY_mTm;
            dvec.gradientZ_mTm = P.hvec.gradientZ_mTm;
            P.uvec.copy_from_device(dvec);
        }
        size_t grid = (S + 255) / 256;
        SWREF_CK(cudaEventRecord(e0));
        sim::cu_sim<<<grid, 256, 0>>>(P.param, P.uvec, fm.empty() ? nullptr : thrust::raw_pointer_cast(d_fm.data()),
                                      thrust::raw_pointer_cast(d_mask.data()), thrust::raw_pointer_cast(d_M0.data()),
                                      thrust::raw_pointer_cast(d_XYZ0.data()),
                                      thrust::raw_pointer_cast(d_M1.data()) + 3 * nTE * S * k,
                                      thrust::raw_pointer_cast(d_XYZ1.data()) + 3 * S * P.trj * k,
                                      thrust::raw_pointer_cast(d_T.data()) + nTE * S * k);
        SWREF_CK(cudaEventRecord(e1));
        SWREF_CK(cudaGetLastError());
        SWREF_CK(cudaDeviceSynchronize());
        float ms = 0.f;
        SWREF_CK(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    thrust::copy(d_M1.begin(), d_M1.end(), M1);
    thrust::copy(d_XYZ1.begin(), d_XYZ1.end(), XYZ1);
    thrust::copy(d_T.begin(), d_T.end(), T);
    if (kernel_ms) *kernel_ms = total;
    return 0;
}
#endif

} /* extern "C" */
