// spinwalk_b200/csrc/walk_fast.cuh — SWK_MODE_FAST walk kernel (the product path), sm_100a.
//
// Same stochastic process as the reference's time loop (src/sim/kernels.cu:107-232),
// engineered for the B200 issue pipes instead of being a translation of it:
//   * position = one 32-bit FIXED-POINT word per axis in grid units: voxel index in the high bits, FB fraction
//     bits below.  A step is  pos += int(n * sigma)  done as one FFMA (magic-number rounding) + one IADD3;
//     "did the voxel change" is (old ^ new) >> FB, so the common no-change step touches neither the index
//     arithmetic nor memory, and there is no float<->int conversion (quarter-rate pipe) anywhere in the loop;
//   * Philox4x32-10 with a fixed key: the ten round keys are immediates of the LOP3s (2 IMAD.WIDE + 2 LOP3 per
//     round), Box-Muller on the MUFU pipe (lg2 / sqrt / sin / cos approx) — the random numbers of attempt n+1
//     are generated between ISSUING the voxel gather of attempt n and CONSUMING it, so the gather latency
//     (L2 ~250 cyc, HBM ~600+ cyc) overlaps ~65 independent instructions per warp;
//   * one 4-byte gather per voxel change: the packed voxel word (FP32 field | 4-bit substrate id), read-only path;
//   * per-thread time: lanes of a warp re-converge only at sequence events, so a lane that has to redraw
//     (permeability rejection, kernels.cu:154-160) does not stall the other 31 per step;
//   * 32-bit voxel indices (V < 2^32); registers kept low enough for >= 4 CTAs (32 warps) per SM.
#pragma once

#include "walk_kernel.cuh"

#ifndef SWK_FAST_MIN_BLOCKS
#define SWK_FAST_MIN_BLOCKS 5 // 48 registers: 40 warps per SM (measured best on the C2 mix of scales, profiles/)
#endif

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

// Philox2x32-10 (same paper) for the permeability uniform (kernels.cu:154): one 32-bit word is needed per test, and the test runs
// whenever ANY lane of the warp changes substrate with 0 < P < 1 — half the multiplies of a 4x32 block (10 IMAD.WIDE + 10 LOP3).
//   counter = (attempt index of the walker, global spin id); key = a 32-bit fold of the run's seed (engine.cu), whose ten round keys
//   key + r * 0x9E3779B9 arrive as kernel parameters, i.e. as constant-bank operands of the LOP3s.
// A different generator AND key than the displacement stream: the two are independent (the reference draws both from copies of
// one minstd stream).
__device__ __forceinline__ uint32_t philox2x32_10(uint32_t c0, uint32_t c1, const uint32_t (&key)[10])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
        c0 = hi ^ key[r] ^ c1;
        c1 = lo;
    }
    return c0;
}

// three N(0,1) from 128 random bits: Box-Muller, 23-bit uniforms, hardware transcendental approximations.
// |n| <= sqrt(2 ln 2^23) = 5.65 by construction (what bounds the fixed-point step below).
__device__ __forceinline__ void normals3_fast(const uint4 r, float &n0, float &n1, float &n2)
{
    const float kNeg2Ln2 = -1.3862943611198906f, k2Pi = 6.283185307179586f;
    const float ua = 2.0f - __uint_as_float((r.x >> 9) | 0x3f800000u); // (0,1]
    const float ub = 2.0f - __uint_as_float((r.z >> 9) | 0x3f800000u);
    const float ta = fmaf(__uint_as_float((r.y >> 9) | 0x3f800000u), k2Pi, -k2Pi); // [0, 2 pi)
    const float tb = fmaf(__uint_as_float((r.w >> 9) | 0x3f800000u), k2Pi, -k2Pi);
    const float ra = mufu_sqrt(kNeg2Ln2 * mufu_lg2(ua));
    const float rb = mufu_sqrt(kNeg2Ln2 * mufu_lg2(ub));
    n0 = ra * mufu_cos(ta);
    n1 = ra * mufu_sin(ta);
    n2 = rb * mufu_cos(tb);
}

// six N(0,1) from ONE 128-bit Philox block: three Box-Muller pairs.  Pair i takes its radius uniform from the top 23 bits
// of word i (r.x / r.y / r.z) and its 19-bit angle uniform from the remaining 9 bits of that word followed by a 10-bit field
// of r.w — 126 of the 128 bits, no bit used twice.  |n| <= sqrt(2 ln 2^23) = 5.65.  One block feeds TWO attempts of the walk.
//   `one` holds 0x3f800000 in a REGISTER (it arrives as a kernel argument so that ptxas cannot turn it back into an
//   immediate): (x & 0x007ffff0) | one is then a single LOP3 instead of two.
__device__ __forceinline__ uint32_t and_or(const uint32_t x, const uint32_t one)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, 0x007ffff0, %2, 0xEA;" : "=r"(d) : "r"(x), "r"(one)); // (x & imm) | one
    return d;
}
__device__ __forceinline__ void bm_pair(const uint32_t w, const uint32_t wlo, const uint32_t one, float &a, float &b)
{
    const float kNeg2Ln2 = -1.3862943611198906f, k2Pi = 6.283185307179586f;
    const float u = 2.0f - __uint_as_float((w >> 9) + 0x3f800000u);                       // (0,1], 23 bits: w[31:9]   (LEA.HI)
    const float t = __uint_as_float(and_or(__funnelshift_l(wlo, w, 14), one)) * k2Pi;     // [2 pi, 4 pi): w[8:0] ++ wlo[31:22]
    const float r = mufu_sqrt(kNeg2Ln2 * mufu_lg2(u));
    a = r * mufu_cos(t);
    b = r * mufu_sin(t);
}
__device__ __forceinline__ void normals6_fast(const uint4 r, const uint32_t one, float &a0, float &a1, float &a2, float &b0, float &b1, float &b2)
{
    bm_pair(r.x, r.w, one, a0, a1);       // angle bits: r.x[8:0] ++ r.w[31:22]
    bm_pair(r.y, r.w << 10, one, a2, b0); //             r.y[8:0] ++ r.w[21:12]
    bm_pair(r.z, r.w << 20, one, b1, b2); //             r.z[8:0] ++ r.w[11:2]
}

// ---- fixed-point grid coordinates -------------------------------------------------------------------------------
// pos = voxel << FB | fraction, FB chosen per launch-block (per scale) on the host side of the kernel:
//   (a) 5.65 sigma_vox 2^FB < 2^22   so that the magic-number rounding of the step is exact to one unit,
//   (b) (n + 1) 2^FB + 2^22 <= 2^32  so that a step across the far wall cannot wrap to a valid position;
// a step below 0 wraps to >= 2^32 - 2^22, which (b) keeps above every valid position: both walls are caught by ONE
// unsigned compare of the voxel index against n.
constexpr float kMagic = 12582912.0f;           // 1.5 * 2^23: float(x + kMagic) holds round(x) in its low mantissa bits
constexpr uint32_t kMagicBits = 0x4B400000u;

__device__ __forceinline__ uint32_t fx_step(uint32_t pos, float n, float sg)
{
    return pos + (uint32_t)__float_as_int(fmaf(n, sg, kMagic)) - kMagicBits;
}

// FoV boundary of one axis (rare; out of line).  kernels.cu:133-136.  `q` is the tentative position, already outside [0, n).
__device__ __noinline__ uint32_t fov_boundary(uint32_t q, const uint32_t p, const uint32_t n, const uint32_t fb, const int cross)
{
    const uint32_t span = n << fb;
    if (cross) { // periodic: re-enter from the other side (single wrap, like the reference)
        q = ((int32_t)(q - p) < 0) ? q + span : q - span;
    } else {     // the reference reverses the step: new = old - rnd
        q = p - (q - p);
    }
    if ((q >> fb) >= n) q = p; // |step| exceeds the distance to both walls: stay
    return q;
}

// The voxel gather.  A plain ld.global.nc makes L2 fetch the whole 128 B line from HBM (measured: 3.7 sectors per
// missed sector, tools/gather_probe.cu); the L2::64B prefetch-size qualifier (LDG.E.LTC64B) halves that traffic at the
// same gather rate — the rate is bound by HBM row activations, not bytes — and keeps the board under its power cap.
// (An L2 evict_first policy for the blocks of the small FoV scales, which touch the whole table at random, was measured
// on C2 and made the pass 2-7 % slower at every threshold; the gather therefore carries no eviction hint.)
__device__ __forceinline__ uint32_t ldg_voxel(const uint32_t *p)
{
    uint32_t v;
    asm("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// VOX selects how a voxel is fetched: 0 = mask only (no fieldmap), 1 = mask byte + FP32 field (two gathers issued
// together), 2 = one packed 32-bit word (field with its 4 low mantissa bits replaced by the substrate id).
// 3 = the packed word of a phantom that is invariant along z (every cylinder phantom), fetched from its [nx][ny] slab: the same
// words as variant 2 from a table nz times smaller (opt-in: SWK_RUN_ZSLAB, engine.cu run_impl).
enum { VOX_MASK = 0, VOX_SPLIT = 1, VOX_PACKED = 2, VOX_SLAB = 3 };

// GRUNS: the sequence holds runs of gradient samples (taken inside the inner loop); sequences without them get a kernel without that code.
template <bool STATS, bool RECORD, int VOX, bool GRUNS>
__global__ void __launch_bounds__(kBlock, SWK_FAST_MIN_BLOCKS) walk_fast_kernel(const __grid_constant__ WalkArgs A)
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
    // per-substrate step sigma in fixed-point grid units for this block's scale: sgt[sub][axis]
    float *sgt = bsum + n_bsum;
    for (uint32_t i = threadIdx.x; i < n_bsum; i += kBlock) bsum[i] = 0.f;

    const uint32_t k = blockIdx.x % A.n_scales;
    const float scale = __ldg(A.scales + k);
    float fscale = 1.f, gscale = 1.f, lin_pc = A.lin_pc;
    if (A.scale_type == SWK_SCALE_FOV) fscale = scale;
    else if (A.scale_type == SWK_SCALE_GRADIENT) gscale = scale;
    else if (A.scale_type == SWK_SCALE_PHASE_CYCLING) lin_pc = __fmul_rn(A.lin_pc, scale); // monte_carlo.cu:303

    const uint32_t n3[3] = {A.nx, A.ny, A.nz};
    double inv_h[3]; // grid units per metre at scale 1
#pragma unroll
    for (int i = 0; i < 3; i++) inv_h[i] = (double)n3[i] / (double)A.fov[i];

    // ---- fraction bits of this block's scale (block-uniform) ----
    uint32_t fb;
    {
        const double *tsig = blob_ptr<double>(A.blob, L.sigma);
        double smax = 0.;
        for (uint32_t s = 0; s < L.n_sub; s++)
            for (int i = 0; i < 3; i++) smax = fmax(smax, tsig[s] * inv_h[i] / (double)fscale);
        const uint32_t nmax = max(n3[0], max(n3[1], n3[2]));
        int f = 22;
        while (f > 0 && ((double)(nmax + 1u) * (double)(1u << f) + 4194304. > 4294967296.)) f--; // (b)
        while (f > 0 && 5.7 * smax * (double)(1u << f) >= 4194304.) f--;                           // (a)
        fb = (uint32_t)f;
        for (uint32_t i = threadIdx.x; i < 3u * L.n_sub; i += kBlock) {
            const uint32_t ax = i % 3u;
            const double ih = ax == 0 ? inv_h[0] : (ax == 1 ? inv_h[1] : inv_h[2]);
            sgt[i] = (float)(tsig[i / 3u] * ih / (double)fscale * (double)(1u << f));
        }
        if (threadIdx.x < 3u) { // metres per fixed-point unit x (1e-3 dt 1e-6 gamma 180/pi), kernels.cu:185
            const double ih = threadIdx.x == 0 ? inv_h[0] : (threadIdx.x == 1 ? inv_h[1] : inv_h[2]);
            sgt[3u * L.n_sub + threadIdx.x] = (float)((double)fscale / (ih * (double)(1u << f)) * 1e-3 * (double)A.timestep_us * 1e-6 * kGamma * kRad2Deg);
        }
    }
    __syncthreads();
    // metres per fixed-point unit at this scale (events and outputs only)
    const double unit_m[3] = {(double)fscale / (inv_h[0] * (double)(1u << fb)), (double)fscale / (inv_h[1] * (double)(1u << fb)),
                              (double)fscale / (inv_h[2] * (double)(1u << fb))};

    const int32_t  *tl_time = blob_ptr<int32_t>(B, L.tl_time);
    const uint32_t *tl_mask = blob_ptr<uint32_t>(B, L.tl_mask), *tl_run = blob_ptr<uint32_t>(B, L.tl_run);
    const float *gtx = blob_ptr<float>(B, L.gx), *gty = blob_ptr<float>(B, L.gy), *gtz = blob_ptr<float>(B, L.gz);
    const float *umk = sgt + 3u * L.n_sub; // degrees of phase per (mT/m x fixed-point unit) and axis, for gradient runs
    const float *tT1 = blob_ptr<float>(B, L.T1s), *tT2 = blob_ptr<float>(B, L.T2s), *tpXY = blob_ptr<float>(B, L.pXY);

    // ---- which spin ----
    const uint32_t j = A.j_first + (blockIdx.x / A.n_scales) * kBlock + threadIdx.x;
    bool alive = j < A.j_end;
    const uint32_t jl = alive ? (A.order ? __ldg(A.order + (A.order_per_scale ? (size_t)k * A.n_local : 0) + j) : j) : 0u;
    const uint32_t spin_no = A.spin_first + jl; // GLOBAL spin id: RNG key and dephasing term

    float m[3] = {0.f, 0.f, 1.f};
    uint32_t p0, p1, p2; // fixed-point position
    {
        uint32_t pp[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float x0 = 0.f;
            if (alive) {
                x0 = __ldg(A.xyz0 + 3 * (size_t)jl + i);
                if (A.m0) m[i] = __ldg(A.m0 + 3 * (size_t)jl + i);
            }
            double g = (double)x0 * inv_h[i] * (double)(1u << fb);
            const double hi = (double)n3[i] * (double)(1u << fb) - 1.; // spins exactly on the far wall start in the last voxel
            g = fmin(fmax(g, 0.), hi);
            pp[i] = (uint32_t)g;
        }
        p0 = pp[0]; p1 = pp[1]; p2 = pp[2];
    }
    // ---- resuming a long run after a re-binning pause (engine.cu: run_impl): position, magnetisation, substrate and RNG
    //      block counter come back from the state arrays; everything else of the per-TR state is reset at a TR start anyway ----
    uint32_t blk = 0;
    const size_t st_idx = (size_t)k * A.n_local + jl;
    if (A.scan_first > 0 && alive) {
        const uint4 sa = A.state_a[st_idx], sb = A.state_b[st_idx];
        p0 = sa.x; p1 = sa.y; p2 = sa.z; blk = sa.w;
        m[0] = __uint_as_float(sb.x); m[1] = __uint_as_float(sb.y); m[2] = __uint_as_float(sb.z);
    }
    const uint32_t ny = A.ny, nz = A.nz;
    uint32_t ind_cur = ((p0 >> fb) * ny + (p1 >> fb)) * nz + (p2 >> fb);
    uint32_t ts_old = alive ? (uint32_t)__ldg(A.mask + ind_cur) : 0u;
    if (A.scan_first > 0 && alive) {
        const uint32_t meta = A.state_b[st_idx].w; // substrate | lost << 8
        ts_old = meta & 0xffu;
        if (meta & 0x100u) alive = false; // lost in an earlier launch (already counted there)
    }
    const bool has_field = VOX != VOX_MASK;
    const float field_k = A.field_k;
    // The reference loads field / T1 / T2 at the first accepted step (kernels.cu:91,150-170).  Holding the field of
    // the CURRENT voxel from the start is equivalent: a first step that stays in the voxel reads this very value.
    float field = 0.f;
    if (alive && VOX == VOX_SPLIT) field = __fmul_rn(__ldg(A.fieldmap + ind_cur), field_k);
    if (alive && VOX == VOX_PACKED) field = __fmul_rn(__uint_as_float(__ldg(A.packed + ind_cur) & 0xfffffff0u), field_k);
    if (alive && VOX == VOX_SLAB) field = __fmul_rn(__uint_as_float(__ldg(A.packed + ((p0 >> fb) * ny + (p1 >> fb))) & 0xfffffff0u), field_k);
    float sg0 = sgt[3 * ts_old], sg1 = sgt[3 * ts_old + 1], sg2 = sgt[3 * ts_old + 2];

    uint32_t itr = 0;
    const uint32_t seed_lo = (uint32_t)A.seed;
    const uint32_t seed_hi_walk = ((uint32_t)(A.seed >> 32) & 0x3fffffffu) | (STREAM_WALK << 30);
    uint32_t st_mask = 0, st_field = 0, st_rej = 0, st_steps = 0; // per thread and launch: < 2^32
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
    // One Philox block feeds TWO attempts: the even attempt of block `blk` steps by (na*), the odd one by (nb*).
    const uint32_t kOne = A.one_bits; // 0x3f800000, deliberately opaque to ptxas (see and_or)
    float na0, na1, na2, nb0, nb1, nb2;
    normals6_fast(philox_fixed(blk, seed_lo, spin_no, seed_hi_walk), kOne, na0, na1, na2, nb0, nb1, nb2);

    for (uint32_t scan = A.scan_first; scan < A.scan_end; scan++) {
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

        // The timeline is walked entry by entry: steps up to and including the entry's timepoint, then its events.  A RUN of
        // gradient-only samples at consecutive timepoints (a PGSE lobe: one sample per step for 10 ms) is taken in one go
        // instead: the plain steps before its first timepoint (part 0), then one step + one gradient sample per timepoint inside
        // the inner loop itself (part 1) — no segment restart per sample.
        for (uint32_t ev = 0; ev <= L.n_tl;) {
            const uint32_t ev_time = ev < L.n_tl ? (uint32_t)tl_time[ev] : n_tp;
            const uint32_t run = ev < L.n_tl ? tl_run[ev] : 0u;
            const bool is_run = GRUNS && run >= 2u && ev_time < n_tp;
            const uint32_t run_len = is_run ? min(run, n_tp - ev_time) : 0u;
          for (int part = 0; part < (is_run ? 2 : 1); part++) {
            const bool grun = GRUNS && part == 1;
            const uint32_t t_stop = is_run ? (grun ? ev_time + run_len : ev_time) : (ev_time < n_tp ? ev_time + 1u : n_tp);
            const uint32_t grad_first = cnt_grad;
            int rem = alive ? (int)(t_stop - t) : 0; // accepted steps still to take in this part

            // =============================== inner loop ===============================
            // one attempt = one tentative step (kernels.cu:130-170).  `overlap` is independent work (random numbers of later
            // attempts) placed between ISSUING the voxel gather and CONSUMING it.  Returns true when the step was accepted.
            auto attempt = [&](const float a0, const float a1, const float a2, const uint32_t perm_ctr, auto &&overlap) -> bool {
                uint32_t q0 = fx_step(p0, a0, sg0), q1 = fx_step(p1, a1, sg1), q2 = fx_step(p2, a2, sg2);
                const bool hop = (((p0 ^ q0) | (p1 ^ q1) | (p2 ^ q2)) >> fb) != 0u;
                uint32_t ts = ts_old;
                float fv = 0.f;
                uint32_t ind_new = ind_cur; // STATS bookkeeping only
                bool chg = false;
                if (hop) {
                    uint32_t v0 = q0 >> fb, v1 = q1 >> fb, v2 = q2 >> fb;
                    if ((v0 >= n3[0]) | (v1 >= n3[1]) | (v2 >= n3[2])) { // FoV boundary (kernels.cu:133-136), rare
                        if (v0 >= n3[0]) { q0 = fov_boundary(q0, p0, n3[0], fb, A.cross_fov); v0 = q0 >> fb; }
                        if (v1 >= n3[1]) { q1 = fov_boundary(q1, p1, n3[1], fb, A.cross_fov); v1 = q1 >> fb; }
                        if (v2 >= n3[2]) { q2 = fov_boundary(q2, p2, n3[2], fb, A.cross_fov); v2 = q2 >> fb; }
                    }
                    ind_new = (v0 * ny + v1) * nz + v2;
                    if (STATS) { chg = (ind_new != ind_cur) | fresh; st_mask += chg; }
                    if (VOX == VOX_PACKED || VOX == VOX_SLAB) { // one gather
                        const uint32_t w = ldg_voxel(A.packed + (VOX == VOX_SLAB ? v0 * ny + v1 : ind_new));
                        ts = w & 15u;
                        fv = __uint_as_float(w & 0xfffffff0u);
                    } else {                 // both gathers issued back to back
                        ts = __ldg(A.mask + ind_new);
                        if (VOX == VOX_SPLIT) fv = __ldg(A.fieldmap + ind_new);
                    }
                } else if (STATS && fresh) {
                    st_mask++; st_field++;
                }
                overlap();
                if (hop) { // kernels.cu:150-170
                    if (ts != ts_old) {
                        // accept iff u < P_XY[from][to], u in [0,1) (kernels.cu:154).  P <= 0 always rejects and P >= 1 always accepts:
                        // the uniform (its own Philox2x32 stream, so skipping a draw changes nothing else) is only generated in between.
                        const float pxy = tpXY[ts_old * L.n_sub + ts];
                        bool reject = pxy <= 0.f;
                        if (pxy > 0.f && pxy < 1.f) reject = u01_open1(philox2x32_10(perm_ctr, spin_no, A.perm_key)) >= pxy;
                        if (reject) {
                            if (STATS) st_rej++;
                            if (itr++ > A.max_iter) { alive = false; lost = true; rem = 0; }
                            return false; // redraw from the old position; time does not advance
                        }
                        ts_old = ts;
                        sg0 = sgt[3 * ts]; sg1 = sgt[3 * ts + 1]; sg2 = sgt[3 * ts + 2];
                    }
                    if (has_field) field = __fmul_rn(fv, field_k); // monte_carlo.cu:244
                    if (STATS) { st_field += chg; ind_cur = ind_new; }
                }
                if (STATS) { fresh = false; st_steps++; }
                p0 = q0; p1 = q1; p2 = q2;
                acc += field; // kernels.cu:171-172
                itr = 0;
                if (GRUNS && grun) { // gradient sample of this timepoint, at the NEW position (kernels.cu:181-187); FP32 here, FP64 in the event path
                    const float gx = __fmul_rn(gtx[cnt_grad], gscale), gy = __fmul_rn(gty[cnt_grad], gscale), gz = __fmul_rn(gtz[cnt_grad], gscale);
                    acc += gx * ((float)p0 * umk[0]) + gy * ((float)p1 * umk[1]) + gz * ((float)p2 * umk[2]);
                    cnt_grad++;
                }
                if (RECORD) { // kernels.cu:218-221 (diagnostic mode)
                    if (X1) {
                        float *slot = X1 + 3 * ((size_t)scan * n_tp + (t_stop - (uint32_t)rem));
                        slot[0] = (float)((double)p0 * unit_m[0]); slot[1] = (float)((double)p1 * unit_m[1]); slot[2] = (float)((double)p2 * unit_m[2]);
                    }
                }
                return true;
            };
            // The integer half of the next block (Philox rounds) overlaps the gather of the even attempt, the float half
            // (Box-Muller) that of the odd attempt.  A segment always starts on a fresh block (lanes of a warp stay in phase).
            while (rem > 0) {
                uint4 raw;
                if (attempt(na0, na1, na2, 2u * blk, [&] { raw = philox_fixed(blk + 1u, seed_lo, spin_no, seed_hi_walk); })) rem--;
                if (rem <= 0) { // the segment ends on an even attempt: (nb*) of this block are dropped
                    normals6_fast(raw, kOne, na0, na1, na2, nb0, nb1, nb2);
                    blk++;
                    break;
                }
                if (attempt(nb0, nb1, nb2, 2u * blk + 1u, [&] { normals6_fast(raw, kOne, na0, na1, na2, nb0, nb1, nb2); })) rem--;
                blk++;
            }
            t = t_stop - (uint32_t)rem;
            if (grun) cnt_grad = grad_first + run_len; // also for lanes that are no longer alive
          }
            // ============================ end of inner loop ============================
            if (is_run) { ev += run_len; continue; } // (a run clipped by the end of the TR is followed by entries >= n_tp only)
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
                    const double X = (double)p0 * unit_m[0], Y = (double)p1 * unit_m[1], Z = (double)p2 * unit_m[2];
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
            ev++;
        }
        if (alive) { // end of TR (kernels.cu:226-231)
            const float dt_s = (float)((double)((n_tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
            dephase_relax(m, acc, tT1[ts_old], tT2[ts_old], dt_s);
        }
    }

    // ---- final position (kernels.cu:220-221 leaves the last committed position in xyz1) ----
    if (A.scan_end < A.n_scans) { // pause at a TR boundary: (na*, nb*) are the untouched normals of block `blk`, so the counter is the whole RNG state
        if (j < A.j_end) {
            A.state_a[st_idx] = make_uint4(p0, p1, p2, blk);
            A.state_b[st_idx] = make_uint4(__float_as_uint(m[0]), __float_as_uint(m[1]), __float_as_uint(m[2]), ts_old | ((alive ? 0u : 1u) << 8));
            A.state_vox[st_idx] = ((p0 >> fb) * ny + (p1 >> fb)) * nz + (p2 >> fb);
        }
    }
    if (!RECORD && X1 && j < A.j_end && A.scan_end == A.n_scans) {
        X1[0] = (float)((double)p0 * unit_m[0]); X1[1] = (float)((double)p1 * unit_m[1]); X1[2] = (float)((double)p2 * unit_m[2]);
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
            unsigned long long c0 = st_steps, c1 = st_mask, c2 = st_field, c3 = st_rej;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                c2 += __shfl_xor_sync(0xffffffffu, c2, o);
                c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            }
            if ((threadIdx.x & 31u) == 0) {
                atomicAdd(A.counters + 0, c0);
                atomicAdd(A.counters + 1, c1);
                atomicAdd(A.counters + 2, c2);
                atomicAdd(A.counters + 3, c3);
            }
        }
        const unsigned lost_w = __popc(__ballot_sync(0xffffffffu, lost));
        if ((threadIdx.x & 31u) == 0 && lost_w) atomicAdd(A.counters + 4, (unsigned long long)lost_w);
    }
}

} // namespace swk
