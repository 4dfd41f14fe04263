// spinwalk_b200/csrc/walk_fast.cuh — SWK_MODE_FAST walk kernel (the product path), sm_100a.
//
// Same stochastic process as the reference's time loop (src/sim/kernels.cu:107-232),
// engineered for the B200 issue pipes instead of being a translation of it:
//   * position = (voxel index, FP32 fraction of a voxel) per axis; a step is 3 FFMA; "did the voxel
//     change" is three unsigned compares on the fraction bits (fraction still in [0,1) <=> no change),
//     so the common no-change step touches neither the index arithmetic nor memory;
//   * Philox4x32-10 with the ten round keys precomputed on the host and read as constant-bank
//     operands (40 integer instructions per 128 random bits), Box-Muller on the MUFU pipe
//     (lg2 / sqrt / sin / cos approx) — the random numbers for attempt n+1 are generated between
//     ISSUING the mask/field gathers of attempt n and CONSUMING them, so the dependent-gather
//     latency (L2 ~250 cyc, HBM ~600+ cyc) overlaps ~70 independent instructions per warp;
//   * mask and field gathers of one voxel change are issued back to back (the reference's are
//     dependent: mask -> permeability test -> field), through the read-only path;
//   * per-thread time t: lanes of a warp re-converge only at sequence events, so a lane that has
//     to redraw (permeability rejection, kernels.cu:154-160) does not stall the other 31 per step;
//   * 32-bit voxel indices (V < 2^32), <= 64 registers => 4 CTAs (32 warps) per SM.
#pragma once

#include "walk_kernel.cuh"

namespace swk {

__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Philox4x32-10 (Salmon et al., SC'11) with a FIXED key so that the ten round keys are immediates of the
// LOP3s (2 IMAD.WIDE + 2 LOP3 per round, no key registers).  The run's seed lives in the counter instead:
//   counter = (attempt counter, seed[31:0], global spin id, stream tag << 30 | seed[61:32])
// Philox is a bijection of the counter for any key, so distinct (seed, spin, attempt, stream) tuples
// give distinct, decorrelated 128-bit blocks.  Every scale replays the same stream, like the reference
// re-seeding seed+spin for each scale (kernels.cu:77-88).
__device__ __forceinline__ uint4 philox_fixed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    constexpr uint32_t K0 = 0x243F6A88u, K1 = 0x85A308D3u; // pi
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
        const uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
        c0 = h1 ^ c1 ^ (K0 + (uint32_t)r * W0);
        c2 = h0 ^ c3 ^ (K1 + (uint32_t)r * W1);
        c1 = l1;
        c3 = l0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// three N(0,1) from 128 random bits: Box-Muller, 23-bit uniforms, hardware transcendental approximations
__device__ __forceinline__ void normals3_fast(const uint4 r, float &n0, float &n1, float &n2)
{
    const float kNeg2Ln2 = -1.3862943611198906f, k2Pi = 6.283185307179586f;
    const float ua = 2.0f - __uint_as_float((r.x >> 9) | 0x3f800000u); // (0,1]
    const float ub = 2.0f - __uint_as_float((r.z >> 9) | 0x3f800000u);
    const float ta = __uint_as_float((r.y >> 9) | 0x3f800000u) - 1.0f; // [0,1)
    const float tb = __uint_as_float((r.w >> 9) | 0x3f800000u) - 1.0f;
    const float ra = mufu_sqrt(kNeg2Ln2 * mufu_lg2(ua));
    const float rb = mufu_sqrt(kNeg2Ln2 * mufu_lg2(ub));
    n0 = ra * mufu_cos(k2Pi * ta);
    n1 = ra * mufu_sin(k2Pi * ta);
    n2 = rb * mufu_cos(k2Pi * tb);
}

// FoV boundary of one axis (rare; out of line, arguments and result in registers).  kernels.cu:133-136
struct FracVox { float g; int v; };
__device__ __noinline__ FracVox fov_boundary(float g, int v, const float pf, const float d, const int pv, const int n, const int cross)
{
    if (cross) { // periodic: re-enter from the other side
        v %= n;
        if (v < 0) v += n;
    } else { // the reference reverses the step: new = old - rnd
        const float h = pf - d;
        const int k = __float2int_rd(h);
        g = h - (float)k;
        v = pv + k;
        if ((unsigned)v >= (unsigned)n) { g = pf; v = pv; } // |step| exceeds the distance to both walls: stay
    }
    return FracVox{g, v};
}

// one axis of a voxel change: g = fraction + step lies outside [0,1).  (A fraction that rounds to exactly 1.0f is
// kept as is: it is flagged as a change again on the next step and resolves itself.)
__device__ __forceinline__ void hop_axis(float &g, int &v, const float pf, const float d, const int pv, const int n, const int cross)
{
    const int k = __float2int_rd(g);
    g -= (float)k;
    v = pv + k;
    if ((unsigned)v >= (unsigned)n) {
        const FracVox r = fov_boundary(g, v, pf, d, pv, n, cross);
        g = r.g;
        v = r.v;
    }
}

// VOX selects how a voxel is fetched: 0 = mask only (no fieldmap), 1 = mask byte + FP32 field (two gathers issued
// together), 2 = one packed 32-bit word (field with its 4 low mantissa bits replaced by the substrate id).
enum { VOX_MASK = 0, VOX_SPLIT = 1, VOX_PACKED = 2 };

template <bool STATS, bool RECORD, int VOX>
__global__ void __launch_bounds__(kBlock, 4) walk_fast_kernel(const __grid_constant__ WalkArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const BlobLayout &L = A.L;

    // ---- stage the sequence tables in shared memory ----
    const uint8_t *B = A.blob;
    uint32_t smem_used = 0;
    if (A.blob_in_smem) {
        const uint32_t nw = L.bytes / 4;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.blob);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem);
        for (uint32_t i = threadIdx.x; i < nw; i += kBlock) dst[i] = __ldg(src + i);
        B = smem;
        smem_used = L.bytes;
    }
    float *bsum = reinterpret_cast<float *>(smem + smem_used);
    const uint32_t n_bsum = A.sums ? A.n_te * L.n_sub * 4u : 0u;
    // per-substrate step sigma in grid units for this block's scale: sgt[sub][axis]
    float *sgt = bsum + n_bsum;
    for (uint32_t i = threadIdx.x; i < n_bsum; i += kBlock) bsum[i] = 0.f;

    const uint32_t k = blockIdx.x % A.n_scales;
    const float scale = __ldg(A.scales + k);
    float fscale = 1.f, gscale = 1.f, lin_pc = A.lin_pc;
    if (A.scale_type == SWK_SCALE_FOV) fscale = scale;
    else if (A.scale_type == SWK_SCALE_GRADIENT) gscale = scale;
    else if (A.scale_type == SWK_SCALE_PHASE_CYCLING) lin_pc = __fmul_rn(A.lin_pc, scale); // monte_carlo.cu:303

    const int n3[3] = {(int)A.nx, (int)A.ny, (int)A.nz};
    float inv_h[3]; // grid units per metre at scale 1
#pragma unroll
    for (int i = 0; i < 3; i++) inv_h[i] = (float)n3[i] / A.fov[i];
    {
        const double *tsig = blob_ptr<double>(A.blob, L.sigma);
        for (uint32_t i = threadIdx.x; i < 3u * L.n_sub; i += kBlock) {
            const uint32_t ax = i % 3u;
            const float ih = ax == 0 ? inv_h[0] : (ax == 1 ? inv_h[1] : inv_h[2]);
            sgt[i] = (float)(tsig[i / 3u] * (double)ih / (double)fscale);
        }
    }
    __syncthreads();

    const int32_t  *tl_time = blob_ptr<int32_t>(B, L.tl_time);
    const uint32_t *tl_mask = blob_ptr<uint32_t>(B, L.tl_mask);
    const float *tT1 = blob_ptr<float>(B, L.T1s), *tT2 = blob_ptr<float>(B, L.T2s), *tpXY = blob_ptr<float>(B, L.pXY);

    // ---- which spin ----
    const uint32_t j = (blockIdx.x / A.n_scales) * kBlock + threadIdx.x;
    bool alive = j < A.n_local;
    const uint32_t jl = alive ? (A.order ? __ldg(A.order + j) : j) : 0u;
    const uint32_t spin_no = A.spin_first + jl; // GLOBAL spin id: RNG key and dephasing term

    float m[3] = {0.f, 0.f, 1.f};
    float pf[3];
    int pv[3];
    {
        float x0[3] = {0.f, 0.f, 0.f};
        if (alive) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                x0[i] = __ldg(A.xyz0 + 3 * (size_t)jl + i);
                if (A.m0) m[i] = __ldg(A.m0 + 3 * (size_t)jl + i);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float g = x0[i] * inv_h[i];
            int v = __float2int_rd(g);
            v = max(0, min(v, n3[i] - 1));
            pv[i] = v;
            pf[i] = fminf(fmaxf(g - (float)v, 0.f), 0.99999994f); // spins exactly on the far wall start in the last voxel
        }
    }
    const uint32_t ny = A.ny, nz = A.nz;
    uint32_t ind_cur = ((uint32_t)pv[0] * ny + (uint32_t)pv[1]) * nz + (uint32_t)pv[2];
    uint32_t ts_old = alive ? (uint32_t)__ldg(A.mask + ind_cur) : 0u;
    const bool has_field = VOX != VOX_MASK;
    const float field_k = A.field_k;
    // The reference loads field / T1 / T2 at the first accepted step (kernels.cu:91,150-170).  Holding the field of
    // the CURRENT voxel from the start is equivalent: a first step that stays in the voxel reads this very value.
    float field = 0.f;
    if (alive && VOX == VOX_SPLIT) field = __fmul_rn(__ldg(A.fieldmap + ind_cur), field_k);
    if (alive && VOX == VOX_PACKED) field = __fmul_rn(__uint_as_float(__ldg(A.packed + ind_cur) & 0xfffffff0u), field_k);
    float sg[3] = {sgt[3 * ts_old], sgt[3 * ts_old + 1], sgt[3 * ts_old + 2]};

    uint32_t ctr = 0, itr = 0;
    const uint32_t seed_lo = (uint32_t)A.seed;
    const uint32_t seed_hi_walk = ((uint32_t)(A.seed >> 32) & 0x3fffffffu) | (STREAM_WALK << 30);
    const uint32_t seed_hi_perm = ((uint32_t)(A.seed >> 32) & 0x3fffffffu) | (STREAM_PERMEABILITY << 30);
    unsigned long long st_mask = 0, st_field = 0, st_rej = 0, st_steps = 0;
    bool lost = false;

    const size_t out_row = (size_t)k * A.n_local + jl;
    float *M1 = A.M1 ? A.M1 + out_row * A.n_te * 3 : nullptr;
    uint8_t *Tt = A.T ? A.T + out_row * A.n_te : nullptr;
    float *X1 = A.XYZ1 ? A.XYZ1 + out_row * A.trj * 3 : nullptr;
    if (RECORD && X1 && alive) { // slot 0 starts as the (scaled) initial position (kernels.cu:96)
#pragma unroll
        for (int i = 0; i < 3; i++) X1[i] = __fmul_rn(__ldg(A.xyz0 + 3 * (size_t)jl + i), fscale);
    }

    const uint32_t n_tp = A.n_tp;
    float n0, n1, n2; // normals of the NEXT attempt
    normals3_fast(philox_fixed(ctr, seed_lo, spin_no, seed_hi_walk), n0, n1, n2);

    for (uint32_t scan = 0; scan < A.n_scans; scan++) {
        const bool last_scan = (scan + 1 == A.n_scans);
        { // phase cycling + first RF (kernels.cu:110-120)
            float ph = (float)((double)(A.rf_ph0 + (float)scan * lin_pc) + (double)(scan * (scan + 1u)) / 2.0 * (double)A.quad_pc);
            while (ph > 360.0) ph = (float)(ph - 360.0);
            while (ph < 0) ph = (float)(ph + 360.0);
            float r[3];
            xrot_withphase(A.s, A.c, ph, m, r);
            m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
        }
        uint32_t t = 0, t_old = 0;
        uint32_t cur_rf = 1, cur_te = 0, cnt_deph = 0, cnt_grad = 0;
        float acc = 0.f;
        bool fresh = true; // only for the STATS counters (ind_old = matrix_length+1, kernels.cu:123)

        for (uint32_t ev = 0; ev <= L.n_tl; ev++) {
            const uint32_t ev_time = ev < L.n_tl ? (uint32_t)tl_time[ev] : n_tp;
            const uint32_t t_stop = ev_time < n_tp ? ev_time + 1u : n_tp;
            int rem = alive ? (int)(t_stop - t) : 0; // accepted steps still to take before the next event

            // =============================== inner loop ===============================
            while (rem > 0) {
                float g0 = fmaf(n0, sg[0], pf[0]), g1 = fmaf(n1, sg[1], pf[1]), g2 = fmaf(n2, sg[2], pf[2]);
                // fraction still in [0,1)  <=>  bit pattern below 1.0f (negative floats compare above)
                const bool hop = (__float_as_uint(g0) >= 0x3f800000u) | (__float_as_uint(g1) >= 0x3f800000u) |
                                 (__float_as_uint(g2) >= 0x3f800000u);
                int v0 = pv[0], v1 = pv[1], v2 = pv[2];
                uint32_t ind_new = ind_cur, ts = ts_old;
                float fv = 0.f;
                if (hop) {
                    if (__float_as_uint(g0) >= 0x3f800000u) hop_axis(g0, v0, pf[0], n0 * sg[0], pv[0], n3[0], A.cross_fov);
                    if (__float_as_uint(g1) >= 0x3f800000u) hop_axis(g1, v1, pf[1], n1 * sg[1], pv[1], n3[1], A.cross_fov);
                    if (__float_as_uint(g2) >= 0x3f800000u) hop_axis(g2, v2, pf[2], n2 * sg[2], pv[2], n3[2], A.cross_fov);
                    ind_new = ((uint32_t)v0 * ny + (uint32_t)v1) * nz + (uint32_t)v2;
                    if (VOX == VOX_PACKED) { // one gather: consumed after the next RNG block
                        const uint32_t w = __ldg(A.packed + ind_new);
                        ts = w & 15u;
                        fv = __uint_as_float(w & 0xfffffff0u);
                    } else {                 // both gathers issued back to back, consumed after the next RNG block
                        ts = __ldg(A.mask + ind_new);
                        if (VOX == VOX_SPLIT) fv = __ldg(A.fieldmap + ind_new);
                    }
                }
                // ---- random numbers of the next attempt: independent work that overlaps the gathers ----
                const uint32_t ctr_this = ctr++;
                normals3_fast(philox_fixed(ctr, seed_lo, spin_no, seed_hi_walk), n0, n1, n2);

                if (hop) { // kernels.cu:150-170
                    if (STATS) st_mask += (ind_new != ind_cur || fresh);
                    if (ts != ts_old) {
                        const float u = u01_open1(philox_fixed(ctr_this, seed_lo, spin_no, seed_hi_perm).x);
                        if (u >= tpXY[ts_old * L.n_sub + ts]) {
                            if (STATS) st_rej++;
                            if (itr++ > A.max_iter) { alive = false; lost = true; break; }
                            continue; // redraw from the old position; time does not advance
                        }
                        ts_old = ts;
                        sg[0] = sgt[3 * ts]; sg[1] = sgt[3 * ts + 1]; sg[2] = sgt[3 * ts + 2];
                    }
                    if (STATS) { st_field += (ind_new != ind_cur || fresh); }
                    ind_cur = ind_new;
                    pv[0] = v0; pv[1] = v1; pv[2] = v2;
                    if (has_field) field = __fmul_rn(fv, field_k); // monte_carlo.cu:244
                } else if (STATS && fresh) {
                    st_mask++; st_field++;
                }
                if (STATS) { fresh = false; st_steps++; }
                pf[0] = g0; pf[1] = g1; pf[2] = g2;
                acc += field; // kernels.cu:171-172
                itr = 0;
                if (RECORD) { // kernels.cu:218-221 (diagnostic mode)
                    if (X1) {
                        float *slot = X1 + 3 * ((size_t)scan * n_tp + (t_stop - (uint32_t)rem));
#pragma unroll
                        for (int i = 0; i < 3; i++) slot[i] = (float)(((double)pv[i] + (double)pf[i]) / (double)inv_h[i] * (double)fscale);
                    }
                }
                rem--;
            }
            t = t_stop - (uint32_t)rem;
            // ============================ end of inner loop ============================
            if (ev >= L.n_tl) break;
            if (ev_time >= n_tp) break;

            // ---- events of timepoint ev_time, in the reference's order (kernels.cu:175-215) ----
            const uint32_t mask_ev = tl_mask[ev];
            const uint32_t tp = ev_time;
            if (mask_ev & EV_DEPH) { // kernels.cu:175-178
                if (alive) acc += (float)spin_no * blob_ptr<float>(B, L.deph_deg)[cnt_deph] / (float)A.n_spins_global;
                cnt_deph++;
            }
            if (mask_ev & EV_GRAD) { // kernels.cu:181-187
                if (alive) {
                    const float Gx = __fmul_rn(blob_ptr<float>(B, L.gx)[cnt_grad], gscale), Gy = __fmul_rn(blob_ptr<float>(B, L.gy)[cnt_grad], gscale),
                                Gz = __fmul_rn(blob_ptr<float>(B, L.gz)[cnt_grad], gscale); // monte_carlo.cu:288-290
                    const double X = ((double)pv[0] + (double)pf[0]) / (double)inv_h[0] * (double)fscale;
                    const double Y = ((double)pv[1] + (double)pf[1]) / (double)inv_h[1] * (double)fscale;
                    const double Z = ((double)pv[2] + (double)pf[2]) / (double)inv_h[2] * (double)fscale;
                    double g = __fma_rn((double)Gz, Z, __fma_rn((double)Gx, X, __dmul_rn((double)Gy, Y)));
                    g = g * 1e-3 * (double)A.timestep_us * 1e-6 * kGamma;
                    acc = (float)__fma_rn(g, kRad2Deg, (double)acc);
                }
                cnt_grad++;
            }
            if (mask_ev & EV_RF) { // kernels.cu:190-199
                if (alive) {
                    const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                    dephase_relax(m, acc, tT1[ts_old], tT2[ts_old], dt_s);
                    float r[3];
                    xrot_withphase(blob_ptr<float>(B, L.rf_s)[cur_rf], blob_ptr<float>(B, L.rf_c)[cur_rf], blob_ptr<float>(B, L.rf_ph)[cur_rf], m, r);
                    m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
                    acc = 0.f;
                    t_old = tp;
                }
                cur_rf++;
            }
            if ((mask_ev & EV_ECHO) && last_scan) { // kernels.cu:202-215
                if (alive) {
                    const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                    dephase_relax(m, acc, tT1[ts_old], tT2[ts_old], dt_s);
                    if (M1) { M1[3 * cur_te + 0] = m[0]; M1[3 * cur_te + 1] = m[1]; M1[3 * cur_te + 2] = m[2]; }
                    if (Tt) Tt[cur_te] = (uint8_t)ts_old;
                    acc = 0.f;
                    t_old = tp;
                }
                if (A.sums) { // ensemble sums per substrate: warp shuffle, then shared-memory accumulate
                    const uint32_t lane = threadIdx.x & 31u;
                    for (uint32_t sub = 0; sub < L.n_sub; sub++) {
                        const bool mine = alive && ts_old == sub;
                        const unsigned any = __ballot_sync(0xffffffffu, mine);
                        if (!any) continue;
                        const float sx = warp_sum(mine ? m[0] : 0.f), sy = warp_sum(mine ? m[1] : 0.f), sz = warp_sum(mine ? m[2] : 0.f);
                        if (lane == 0) {
                            float *b = bsum + (cur_te * L.n_sub + sub) * 4u;
                            atomicAdd(b + 0, sx); atomicAdd(b + 1, sy); atomicAdd(b + 2, sz);
                            atomicAdd(b + 3, (float)__popc(any));
                        }
                    }
                }
                cur_te++;
            }
        }
        if (alive) { // end of TR (kernels.cu:226-231)
            const float dt_s = (float)((double)((n_tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
            dephase_relax(m, acc, tT1[ts_old], tT2[ts_old], dt_s);
        }
    }

    // ---- final position (kernels.cu:220-221 leaves the last committed position in xyz1) ----
    if (!RECORD && X1 && j < A.n_local) {
#pragma unroll
        for (int i = 0; i < 3; i++) X1[i] = (float)(((double)pv[i] + (double)pf[i]) / (double)inv_h[i] * (double)fscale);
    }

    // ---- flush block sums and counters ----
    __syncthreads();
    if (A.sums) {
        double *gs = A.sums + (size_t)k * n_bsum;
        for (uint32_t i = threadIdx.x; i < n_bsum; i += kBlock) {
            const float v = bsum[i];
            if (v != 0.f) atomicAdd(gs + i, (double)v);
        }
    }
    if (A.counters) {
        if (STATS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                st_steps += __shfl_xor_sync(0xffffffffu, st_steps, o);
                st_mask += __shfl_xor_sync(0xffffffffu, st_mask, o);
                st_field += __shfl_xor_sync(0xffffffffu, st_field, o);
                st_rej += __shfl_xor_sync(0xffffffffu, st_rej, o);
            }
        }
        const unsigned lost_w = __popc(__ballot_sync(0xffffffffu, lost));
        if ((threadIdx.x & 31u) == 0) {
            if (STATS) {
                atomicAdd(A.counters + 0, st_steps);
                atomicAdd(A.counters + 1, st_mask);
                atomicAdd(A.counters + 2, st_field);
                atomicAdd(A.counters + 3, st_rej);
            }
            if (lost_w) atomicAdd(A.counters + 4, (unsigned long long)lost_w);
        }
    }
}

} // namespace swk
