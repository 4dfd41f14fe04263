// spinwalk_b200/csrc/walk_fast.cuh — SWK_MODE_FAST walk kernel (the product path), sm_100a.
//
// Same stochastic process as the reference's time loop (src/sim/kernels.cu:107-232), engineered for the
// B200 issue pipes — the launch is ISSUE bound once the voxel table is cache resident (ncu, profiles/), so the design
// minimises warp-instructions per attempted step:
//   * SHARED RANDOM STREAM.  The reference re-seeds seed+spin for every scale (kernels.cu:77-88): all FoV scales of a spin
//     replay the SAME displacement stream.  A block therefore walks 32 spins x G scales (one warp per scale, one lane per
//     spin) and generates each spin's normals ONCE per block — Philox4x32-10 + Box-Muller on the MUFU pipe, 16 rounds at a
//     time into shared memory, double buffered — instead of once per (spin, scale): ~47 of the ~90 instructions an attempt
//     used to cost are paid once per G walkers.  (Runs with one scale, and legs resumed after a re-binning pause, use the
//     PRIVATE variant: every thread generates its own normals; same numbers, same results.)
//   * ROUND r of spin j uses normals (j, r): Philox block r >> 1, half r & 1.  A walker executes one attempt per round
//     while it has steps left in its current segment (a segment = the steps up to the next sequence event); when the segment
//     is complete it waits for the next round that is a multiple of kSync and runs the event there.  Which rounds a walker
//     uses therefore depends on its own history only: results are independent of warp / block composition (shard, slice,
//     re-binning and sort invariance are bitwise, tested), and skipped normals are independent of everything the walker used.
//   * the attempt itself is straight-line predicated code (no per-lane branches but three rare ones: FoV wall, substrate
//     change, loss): position = one 32-bit FIXED-POINT word per axis (voxel << FB | fraction), a step is one FFMA
//     (magic-number rounding) + one IADD3 per axis, "did the voxel change" is one integer compare of the table index, the
//     voxel fetch is one predicated 4-byte gather of the packed voxel word (FP32 field whose 4 low mantissa bits are the
//     substrate id) from the read-only path, accept / reject is a select;
//   * 32-bit voxel indices (V < 2^32); ensemble sums in integer fixed point (bit-reproducible, order independent).
#pragma once

#include <type_traits>

#include "walk_kernel.cuh"

#ifndef SWK_FAST_MIN_BLOCKS
#define SWK_FAST_MIN_BLOCKS 5 // PRIVATE variant: 48 registers, 40 warps per SM
#endif
#ifndef SWK_FAST_MULTI_MINB
#define SWK_FAST_MULTI_MINB 3  // MULTI variant (one walker per spin for all gradient / phase-cycling scales): 80 registers, 24 warps per SM (C3, blocks per SM 2 / 3 / 4 / 5 / 6: 4.6 / 5.6 / 5.2 / 5.1 / 3.6e12)
#endif
#ifndef SWK_FAST_SHARED_MAXT
#define SWK_FAST_SHARED_MAXT 320 // SHARED variant: at most 10 scales (warps) per block
#endif
#ifndef SWK_FAST_UNROLL
#define SWK_FAST_UNROLL 4   // rounds unrolled in the SHARED loop (8 with prefetch spills ~10 values per attempt at 48 registers; 4: none)
#endif
#ifndef SWK_FAST_SHARED_MINB
#define SWK_FAST_SHARED_MINB 4   // ... and at most 65536 / (4 x 320) = 51 -> 48 registers: 40 warps per SM
#endif

namespace swk {

__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Philox4x32-10 (Salmon et al., SC'11) with a FIXED key so that the ten round keys are immediates of the
// LOP3s (2 IMAD.WIDE + 2 LOP3 per round, no key registers).  The run's seed lives in the counter instead:
//   counter = (block counter, seed[31:0], global spin id, stream tag << 30 | seed[61:32])
// Philox is a bijection of the counter for any key, so distinct (seed, spin, block, stream) tuples
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

// Philox2x32-10 (same paper) for the permeability uniform (kernels.cu:154): one 32-bit word per test.
//   counter = (round index of the walker, global spin id); key = a 32-bit fold of the run's seed (engine.cu), whose ten round keys
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

// six N(0,1) from ONE 128-bit Philox block: three Box-Muller pairs.  Pair i takes its radius uniform from the top 23 bits
// of word i (r.x / r.y / r.z) and its 19-bit angle uniform from the remaining 9 bits of that word followed by a 10-bit field
// of r.w — 126 of the 128 bits, no bit used twice.  |n| <= sqrt(2 ln 2^23) = 5.65.  One block feeds TWO rounds of the walk.
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
// pos = voxel << FB | fraction, FB chosen per scale on the host (engine.cu scale_constants):
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

// FoV boundary of one axis (rare).  kernels.cu:133-136.  `q` is the tentative position, already outside [0, n).
__device__ __forceinline__ uint32_t fov_boundary(uint32_t q, const uint32_t p, const uint32_t n, const uint32_t fb, const int cross)
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
// together), 2 = one packed 32-bit word (the FP32 field rounded to the nearest value whose 4 low mantissa bits spell the
// substrate id; engine.cu pack_word), 3 = the packed word of a phantom that is invariant along z (every cylinder phantom),
// fetched from its [nx][ny] slab: the same words as variant 2 from a table nz times smaller (L1/L2 resident).
enum { VOX_MASK = 0, VOX_SPLIT = 1, VOX_PACKED = 2, VOX_SLAB = 3 };

constexpr int kUnroll = SWK_FAST_UNROLL;
#ifndef SWK_KSYNC
#define SWK_KSYNC 8
#endif
constexpr uint32_t kSync = SWK_KSYNC;   // a walker runs its sequence events at rounds that are multiples of kSync (see the file header)
constexpr uint32_t kBatch = 16; // SHARED variant: rounds of normals generated per barrier (32 spins x 16 rounds = 256 Philox blocks)

// Per-scale constants, computed once per run on the host in double precision (engine.cu scale_constants) and staged in shared
// memory by the blocks that walk the scale.  float sgt[3 * n_sub] follows at byte offset sgt_off: the step sigma per (substrate,
// axis) in fixed-point units.
struct ScaleConst {
    uint32_t fb;            // fraction bits of the fixed-point position
    float fscale, gscale, lin_pc; // FoV scale (monte_carlo.cu:278-280), gradient scale (:288-290), linear phase cycling (:303)
    float umk[3];           // degrees of phase per (mT/m x fixed-point unit) and axis (gradient runs, kernels.cu:185)
    uint32_t sgt_off;       // byte offset of sgt from the start of this record
    double unit_m[3];       // metres per fixed-point unit at this scale (events and outputs)
    double pos_k[3];        // fixed-point units per UNSCALED metre of XYZ0 (positions and FoV scale together)
    double pos_hi[3];       // n 2^FB - 1: spins exactly on the far wall start in the last voxel
};

enum : uint32_t { SEG_START = 0u, SEG_EVENT = 1u, SEG_RUN0 = 2u, SEG_RUN1 = 3u, SEG_TAIL = 4u };

// SHARED variant: normals of rounds [r0, r0 + kBatch) of a block's 32 spins.  Thread t makes Philox block (r0 >> 1) + (t >> 5) of ITS lane's
// spin and stores the two rounds it feeds as float4 (x, y, z step normals, permeability uniform) at buf[round][lane].  Out of line on purpose:
// its ~30 temporaries must not compete with the registers of the round loop.
__device__ __noinline__ void generate_normals(float4 *buf, const uint32_t r0, const uint32_t spin_no, const uint32_t seed_lo, const uint32_t seed_hi_walk,
                                              const uint32_t one_bits, const uint32_t *perm_key /* nullptr: no permeability draws needed */)
{
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i = threadIdx.x; i < 32u * (kBatch / 2u); i += blockDim.x) {
        const uint32_t h = i >> 5; // (i & 31) == lane: blockDim is a multiple of 32
        float a0, a1, a2, b0, b1, b2;
        normals6_fast(philox_fixed((r0 >> 1) + h, seed_lo, spin_no, seed_hi_walk), one_bits, a0, a1, a2, b0, b1, b2);
        float ua = 0.f, ub = 0.f;
        if (perm_key) { // some 0 < P_XY < 1: the permeability uniforms of the two rounds ride along
            uint32_t key[10];
#pragma unroll
            for (int q = 0; q < 10; q++) key[q] = perm_key[q];
            ua = u01_open1(philox2x32_10(r0 + 2u * h, spin_no, key));
            ub = u01_open1(philox2x32_10(r0 + 2u * h + 1u, spin_no, key));
        }
        buf[(2u * h) * 32u + lane] = make_float4(a0, a1, a2, ua);
        buf[(2u * h + 1u) * 32u + lane] = make_float4(b0, b1, b2, ub);
    }
}

// Sequence state of a walker that is only touched at events lives in shared memory (structure of arrays, one word per thread and
// field) so that the registers of the round loop hold nothing but the walk itself.
enum : uint32_t { ES_M0 = 0u, ES_M1, ES_M2, ES_SCAN, ES_EV, ES_SEG, ES_TSTOP, ES_TOLD, ES_RF, ES_TE, ES_DEPH, ES_GFIRST, ES_RUNLEN, ES_FIELDS };

// ======================================= events (out of line) =======================================
// Where a thread's walker lives: scale, thread slot, spin, shared-memory regions.  A pure function of the launch parameters and the
// thread / block indices, so the out-of-line event code recomputes it instead of receiving a context through local memory.
template <bool SHARED, bool MULTI = false>
struct Geo {
    uint32_t k, k_first, k_loc, j, jl, spin_no, nthr, n_bsum;
    bool spin_ok, valid;
    const uint8_t *B;       // sequence tables (shared memory when they fit, else global)
    long long *bsum;        // block sums, [scales of the block][n_bsum]
    uint8_t *sct;           // scale constants of the block's scales
    uint32_t *es;           // this thread's event state, field f at es[f * nthr]
    float4 *nbuf;           // normals (SHARED)
    __device__ __forceinline__ Geo(const WalkArgs &A, uint8_t *smem)
    {
        nthr = blockDim.x;
        uint32_t off = A.blob_in_smem ? A.L.bytes : 0u; // (multiple of 16)
        B = A.blob_in_smem ? smem : A.blob;
        const uint32_t n_grp = SHARED ? A.group : 1u;                       // scales walked by this block (scale constants)
        const uint32_t n_acc = MULTI ? A.n_multi : n_grp;                   // scales this block accumulates sums for (MULTI: all of the run)
        n_bsum = A.sums_fx ? A.n_te * A.L.n_sub * 4u : 0u;
        bsum = reinterpret_cast<long long *>(smem + off);
        off += (n_bsum * n_acc * 8u + 15u) & ~15u;
        sct = smem + off;
        off += A.scale_stride * n_grp;
        es = reinterpret_cast<uint32_t *>(smem + off) + threadIdx.x;
        off += ES_FIELDS * 4u * nthr; // (nthr is a multiple of 32: stays 16-byte aligned)
        nbuf = reinterpret_cast<float4 *>(smem + off);
        if (SHARED) {
            k_first = A.k_lo + (blockIdx.x % A.n_groups) * A.group;
            k_loc = threadIdx.x >> 5;
            k = k_first + k_loc;
            j = A.j_first + (blockIdx.x / A.n_groups) * 32u + (threadIdx.x & 31u);
        } else {
            const uint32_t kn = A.k_hi - A.k_lo;
            k_loc = 0u;
            k = k_first = A.k_lo + blockIdx.x % kn;
            j = A.j_first + (blockIdx.x / kn) * kBlock + threadIdx.x;
        }
        spin_ok = j < A.j_end;
        valid = spin_ok && k < A.k_hi;
        if (k >= A.k_hi) k = k_first; // a padding warp of the last group reads valid constants and walks nothing
        jl = spin_ok ? (A.order ? __ldg(A.order + ((!SHARED && A.order_per_scale) ? (size_t)k * A.n_local : 0) + j) : j) : 0u;
        spin_no = A.spin_first + jl; // GLOBAL spin id: RNG key and dephasing term
    }
    __device__ __forceinline__ const ScaleConst &sc(const WalkArgs &A) const { return *reinterpret_cast<const ScaleConst *>(sct + (size_t)(k - k_first) * A.scale_stride); }
    __device__ __forceinline__ size_t st_idx(const WalkArgs &A) const { return (size_t)k * A.n_local + jl; }
    // staging slot e of this walker: rows are laid out per warp-sized chunk of thread slots, structure of arrays — [scale][chunk][slot e][lane] — so a
    // warp's echo write is 512 contiguous bytes.  Indexed by the thread slot (coalesced; unpack_rows_kernel un-permutes through the inverse order)
    // or, for the legs of a re-binned run whose order changes, by the spin.
    __device__ __forceinline__ uint4 *stage(const WalkArgs &A, uint32_t e) const { return stage(A, e, k); }
    __device__ __forceinline__ uint4 *stage(const WalkArgs &A, uint32_t e, uint32_t kk) const
    {
        const uint32_t row = A.stage_by_slot ? j : jl;
        return A.stage + (((size_t)kk * A.stage_chunks + (row >> 5)) * A.stage_row + e) * 32u + (row & 31u);
    }
};

struct AdvOut {
    float acc;
    int rem;
    uint32_t cnt_grad, flags;
};
struct AdvOutMulti : AdvOut {
    float accg; // MULTI: phase accrued from the UNSCALED gradients since the last event
};
enum : uint32_t { WF_LOST = 1u, WF_DONE = 2u, WF_GRUN = 4u, WF_FRESH = 8u };

// Called at a round that is a multiple of kSync by a walker whose segment is complete (rem == 0): runs the sequence events that are due
// (kernels.cu:175-215, 226-231, 110-126) and sets up the next segment; `r_next` is the first round the walker will use afterwards.
// Out of line on purpose: the event arithmetic (sincos / exp / FP64) must not compete with the registers of the round loop.
template <bool STATS, bool RECORD, int VOX, bool GRUNS, bool SHARED>
__device__ __noinline__ AdvOut advance_walker(const WalkArgs *pA, const uint32_t p0, const uint32_t p1, const uint32_t p2, const uint32_t wcur, float acc,
                                              uint32_t cnt_grad, uint32_t flags, const uint32_t r_next)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const WalkArgs &A = *pA;
    const BlobLayout &L = A.L;
    const Geo<SHARED> g(A, smem);
    const uint8_t *B = g.B;
    const ScaleConst &SC = g.sc(A);
    uint32_t *es = g.es;
    const uint32_t nthr = g.nthr;
    const bool stage = A.stage != nullptr;
    const uint32_t n_tp = A.n_tp;
    const int32_t  *tl_time = blob_ptr<int32_t>(B, L.tl_time);
    const uint32_t *tl_mask = blob_ptr<uint32_t>(B, L.tl_mask), *tl_run = blob_ptr<uint32_t>(B, L.tl_run);
    const float *gtx = blob_ptr<float>(B, L.gx), *gty = blob_ptr<float>(B, L.gy), *gtz = blob_ptr<float>(B, L.gz);
    const float *tT1 = blob_ptr<float>(B, L.T1s), *tT2 = blob_ptr<float>(B, L.T2s);
    const uint32_t ts = (VOX == VOX_PACKED || VOX == VOX_SLAB) ? (wcur & 15u) : wcur; // the substrate does not change during events
    const bool lost = (flags & WF_LOST) != 0u;

    float m[3] = {__uint_as_float(es[ES_M0 * nthr]), __uint_as_float(es[ES_M1 * nthr]), __uint_as_float(es[ES_M2 * nthr])};
    uint32_t scan = es[ES_SCAN * nthr], ev = es[ES_EV * nthr], seg = es[ES_SEG * nthr], t_stop = es[ES_TSTOP * nthr], t_old = es[ES_TOLD * nthr];
    uint32_t cur_rf = es[ES_RF * nthr], cur_te = es[ES_TE * nthr], cnt_deph = es[ES_DEPH * nthr], grad_first = es[ES_GFIRST * nthr], run_len = es[ES_RUNLEN * nthr];
    const float gscale = SC.gscale;
    int rem = 0;
    bool finished = false;
    for (;;) {
        if (lost) { // abandoned (kernels.cu:155-159)
            if (scan + 1 != A.n_scans) cur_te = 0; // no echo of the last scan was written
            finished = true;
            break;
        }
        if (seg == SEG_EVENT) { // events of timepoint tl_time[ev], in the reference's order (kernels.cu:175-215)
            const uint32_t mask_ev = tl_mask[ev];
            const uint32_t tp = (uint32_t)tl_time[ev];
            if (mask_ev & EV_DEPH) { // kernels.cu:175-178
                acc += (float)g.spin_no * blob_ptr<float>(B, L.deph_deg)[cnt_deph] / (float)A.n_spins_global;
                cnt_deph++;
            }
            if (mask_ev & EV_GRAD) { // kernels.cu:181-187
                const float Gx = __fmul_rn(gtx[cnt_grad], gscale), Gy = __fmul_rn(gty[cnt_grad], gscale), Gz = __fmul_rn(gtz[cnt_grad], gscale); // monte_carlo.cu:288-290
                const double X = (double)p0 * SC.unit_m[0], Y = (double)p1 * SC.unit_m[1], Z = (double)p2 * SC.unit_m[2];
                double g = __fma_rn((double)Gz, Z, __fma_rn((double)Gx, X, __dmul_rn((double)Gy, Y)));
                g = g * 1e-3 * (double)A.timestep_us * 1e-6 * kGamma;
                acc = (float)__fma_rn(g, kRad2Deg, (double)acc);
                cnt_grad++;
            }
            if (mask_ev & EV_RF) { // kernels.cu:190-199
                const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                dephase_relax(m, acc, tT1[ts], tT2[ts], dt_s);
                float rr[3];
                xrot_withphase(blob_ptr<float>(B, L.rf_s)[cur_rf], blob_ptr<float>(B, L.rf_c)[cur_rf], blob_ptr<float>(B, L.rf_ph)[cur_rf], m, rr);
                m[0] = rr[0]; m[1] = rr[1]; m[2] = rr[2];
                acc = 0.f;
                t_old = tp;
                cur_rf++;
            }
            if ((mask_ev & EV_ECHO) && scan + 1 == A.n_scans) { // kernels.cu:202-215
                const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                dephase_relax(m, acc, tT1[ts], tT2[ts], dt_s);
                if (stage) *g.stage(A, cur_te) = make_uint4(__float_as_uint(m[0]), __float_as_uint(m[1]), __float_as_uint(m[2]), ts);
                acc = 0.f;
                t_old = tp;
                if (A.sums_fx) echo_sums_add(g.bsum + (size_t)g.k_loc * g.n_bsum, cur_te * L.n_sub + ts, m);
                cur_te++;
            }
            ev++;
        } else if (seg == SEG_RUN0) { // the plain steps before a run of gradient samples are done: now the run itself
            seg = SEG_RUN1;
            flags |= WF_GRUN;
            grad_first = cnt_grad;
            rem = (int)run_len;
            t_stop += run_len;
            break;
        } else if (seg == SEG_RUN1) {
            flags &= ~WF_GRUN;
            cnt_grad = grad_first + run_len;
            ev += run_len;
        } else {
            if (seg == SEG_TAIL) { // end of TR (kernels.cu:226-231)
                // RE-SYNCHRONISATION of multi-TR runs.  Every permeability rejection costs a walker one round, so the lanes of a warp reach their
                // events at different rounds and — over the ~1100 TRs of a bSSFP run — drift apart completely: the event code below would run once
                // per lane instead of once per warp.  TR number i therefore ends no earlier than round (i + 1) x A.tr_period, where the period
                // (engine.cu) is what a walker without rejections needs for a TR plus a slack of 2 + timepoints / 64 rounds: walkers that lose
                // fewer rounds than the slack per TR — nearly all, except behind walls at small FoV scales — stay on the common schedule, and a
                // walker that fell behind catches up by the slack of every TR.  Deterministic per walker (absolute round numbers).
                if (scan + 1 < A.n_scans && r_next < (scan + 1u) * A.tr_period) break; // called again at the next sync round (rem stays 0); also before a re-binning pause
                const float dt_s = (float)((double)((n_tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                dephase_relax(m, acc, tT1[ts], tT2[ts], dt_s);
                scan++;
                if (scan >= A.scan_end) { finished = true; break; }
            }
            { // start of a TR: phase cycling + first RF (kernels.cu:110-126)
                float ph = (float)((double)(A.rf_ph0 + (float)scan * SC.lin_pc) + (double)(scan * (scan + 1u)) / 2.0 * (double)A.quad_pc);
                // the reference wraps by repeated subtraction (kernels.cu:112-113: hundreds of iterations late in a bSSFP run): whole turns in closed form first
                if (ph > 720.f) ph = (float)((double)ph - 360.0 * floor(((double)ph - 360.0) / 360.0));
                if (ph < -360.f) ph = (float)((double)ph + 360.0 * floor(-(double)ph / 360.0));
                while (ph > 360.0) ph = (float)(ph - 360.0);
                while (ph < 0) ph = (float)(ph + 360.0);
                float rr[3];
                xrot_withphase(A.s, A.c, ph, m, rr);
                m[0] = rr[0]; m[1] = rr[1]; m[2] = rr[2];
                t_stop = 0; t_old = 0; ev = 0;
                cur_rf = 1; cur_te = 0; cnt_deph = 0; cnt_grad = 0;
                acc = 0.f;
                if (STATS) flags |= WF_FRESH;
            }
        }
        // ---- next segment: the steps up to and including the next entry's timepoint, or the plain steps before a run of
        //      gradient-only samples at consecutive timepoints (a PGSE lobe: one sample per step), or the rest of the TR ----
        const uint32_t ev_time = ev < L.n_tl ? (uint32_t)tl_time[ev] : n_tp;
        uint32_t stop;
        if (ev >= L.n_tl || ev_time >= n_tp) { stop = n_tp; seg = SEG_TAIL; }
        else if (GRUNS && tl_run[ev] >= 2u) { stop = ev_time; seg = SEG_RUN0; run_len = min(tl_run[ev], n_tp - ev_time); }
        else { stop = ev_time + 1u; seg = SEG_EVENT; }
        rem = (int)(stop - t_stop);
        t_stop = stop;
        if (rem > 0) break;
    }
    if (finished) { // this launch's scans are complete (or the walker was abandoned)
        flags |= WF_DONE;
        rem = 0;
        if (A.scan_end < A.n_scans) { // pause at a TR boundary: the round index is the whole RNG state
            A.state_a[g.st_idx(A)] = make_uint4(p0, p1, p2, r_next);
            A.state_b[g.st_idx(A)] = make_uint4(__float_as_uint(m[0]), __float_as_uint(m[1]), __float_as_uint(m[2]), ts | ((lost ? 1u : 0u) << 8));
            A.state_vox[g.st_idx(A)] = ((p0 >> SC.fb) * A.ny + (p1 >> SC.fb)) * A.nz + (p2 >> SC.fb);
        } else if (stage) {
            // echoes that never fired (beyond the TR, or after the walker was abandoned) read 0, like the reference's zero-initialised
            // outputs (monte_carlo.cu:256,259-260); the final position is the last committed one (kernels.cu:220-221)
            for (uint32_t e = cur_te; e < A.n_te; e++) *g.stage(A, e) = make_uint4(0u, 0u, 0u, 0u);
            if (!RECORD) *g.stage(A, A.n_te) = make_uint4(__float_as_uint((float)((double)p0 * SC.unit_m[0])), __float_as_uint((float)((double)p1 * SC.unit_m[1])),
                                                    __float_as_uint((float)((double)p2 * SC.unit_m[2])), lost ? 1u : 0u);
        }
    } else {
        es[ES_M0 * nthr] = __float_as_uint(m[0]); es[ES_M1 * nthr] = __float_as_uint(m[1]); es[ES_M2 * nthr] = __float_as_uint(m[2]);
        es[ES_SCAN * nthr] = scan; es[ES_EV * nthr] = ev; es[ES_SEG * nthr] = seg; es[ES_TSTOP * nthr] = t_stop; es[ES_TOLD * nthr] = t_old;
        es[ES_RF * nthr] = cur_rf; es[ES_TE * nthr] = cur_te; es[ES_DEPH * nthr] = cnt_deph; es[ES_GFIRST * nthr] = grad_first; es[ES_RUNLEN * nthr] = run_len;
    }
    AdvOut o;
    o.acc = acc; o.rem = rem; o.cnt_grad = cnt_grad; o.flags = flags;
    return o;
}

// advance_walker for MULTI kernels (one walk for all scales, WalkArgs::n_multi).  The same segment machine as advance_walker above — kept as a
// second function so that the code of the tuned single-scale variants stays exactly what it was (ptxas allocates the registers of a kernel and its
// out-of-line callee together: restructuring the callee moved spills into the round loop of the PRIVATE variant).
// MULTI: the walker stands for the same spin at every scale of a run whose scales act on the
// gradients or on the phase cycling (monte_carlo.cu:288-290, 303) — their walks are identical, only the accrued phase differs, and it is linear in
// the gradient scale: phase_k = acc + gscale_k * accg with acc = field + dephasing terms and accg = the gradient term of the UNSCALED gradients.
// Every event is applied to the magnetisation of each scale in turn (A.mstate, coalesced 16-byte slots); the walk itself is paid once.
template <bool STATS, int VOX, bool GRUNS>
__device__ __noinline__ AdvOutMulti advance_walker_multi(const WalkArgs *pA, const uint32_t p0, const uint32_t p1, const uint32_t p2, const uint32_t wcur, float acc,
                                                         float accg, uint32_t cnt_grad, uint32_t flags, const uint32_t r_next)
{
    constexpr bool MULTI = true, SHARED = false, RECORD = false;
    extern __shared__ __align__(16) uint8_t smem[];
    const WalkArgs &A = *pA;
    const BlobLayout &L = A.L;
    const Geo<SHARED, MULTI> g(A, smem);
    const uint8_t *B = g.B;
    const ScaleConst &SC = g.sc(A);
    uint32_t *es = g.es;
    const uint32_t nthr = g.nthr;
    const bool stage = A.stage != nullptr;
    const uint32_t n_tp = A.n_tp;
    const int32_t  *tl_time = blob_ptr<int32_t>(B, L.tl_time);
    const uint32_t *tl_mask = blob_ptr<uint32_t>(B, L.tl_mask), *tl_run = blob_ptr<uint32_t>(B, L.tl_run);
    const float *gtx = blob_ptr<float>(B, L.gx), *gty = blob_ptr<float>(B, L.gy), *gtz = blob_ptr<float>(B, L.gz);
    const float *tT1 = blob_ptr<float>(B, L.T1s), *tT2 = blob_ptr<float>(B, L.T2s);
    const uint32_t ts = (VOX == VOX_PACKED || VOX == VOX_SLAB) ? (wcur & 15u) : wcur; // the substrate does not change during events
    const bool lost = (flags & WF_LOST) != 0u;
    // relaxation over dt: the same factors for every scale (dephase_relax, kernels.cu:45-52)
    const float T1 = tT1[ts], T2 = tT2[ts];
    const bool relax = T1 >= 0 && T2 >= 0;
    float e1 = 0.f, e2 = 0.f;
    auto relax_factors = [&](const float dt_s) { if (relax) { e1 = expf(-dt_s / T1); e2 = expf(-dt_s / T2); } };

    float m[3] = {0.f, 0.f, 0.f};
    if (!MULTI) { m[0] = __uint_as_float(es[ES_M0 * nthr]); m[1] = __uint_as_float(es[ES_M1 * nthr]); m[2] = __uint_as_float(es[ES_M2 * nthr]); }
    // fn(scale, magnetisation, gradient scale, linear phase cycling) for this walker's scale — MULTI: for every scale of the run in turn.
    // `from_first`: every scale still holds the magnetisation the walker started with, which only slot 0 carries (start of the first TR).
    auto each_scale = [&](const bool from_first, auto &&fn) {
        if (MULTI) {
            uint4 *slot0 = A.mstate + (g.j - A.m_first);
            const uint4 v0 = from_first ? *slot0 : make_uint4(0u, 0u, 0u, 0u);
            for (uint32_t kk = 0; kk < A.n_multi; kk++) {
                uint4 *slot = slot0 + (size_t)kk * A.m_rows;
                const uint4 v = from_first ? v0 : *slot;
                float mm[3] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z)};
                const ScaleConst *sk = reinterpret_cast<const ScaleConst *>(A.scale_tab + (size_t)kk * A.scale_stride);
                fn(kk, kk, mm, __ldg(&sk->gscale), __ldg(&sk->lin_pc));
                *slot = make_uint4(__float_as_uint(mm[0]), __float_as_uint(mm[1]), __float_as_uint(mm[2]), 0u);
            }
        } else fn(g.k, g.k_loc, m, SC.gscale, SC.lin_pc);
    };
    uint32_t scan = es[ES_SCAN * nthr], ev = es[ES_EV * nthr], seg = es[ES_SEG * nthr], t_stop = es[ES_TSTOP * nthr], t_old = es[ES_TOLD * nthr];
    uint32_t cur_rf = es[ES_RF * nthr], cur_te = es[ES_TE * nthr], cnt_deph = es[ES_DEPH * nthr], grad_first = es[ES_GFIRST * nthr], run_len = es[ES_RUNLEN * nthr];
    const float gscale = MULTI ? 1.f : SC.gscale;
    int rem = 0;
    bool finished = false;
    for (;;) {
        if (lost) { // abandoned (kernels.cu:155-159)
            if (scan + 1 != A.n_scans) cur_te = 0; // no echo of the last scan was written
            finished = true;
            break;
        }
        if (seg == SEG_EVENT) { // events of timepoint tl_time[ev], in the reference's order (kernels.cu:175-215)
            const uint32_t mask_ev = tl_mask[ev];
            const uint32_t tp = (uint32_t)tl_time[ev];
            if (mask_ev & EV_DEPH) { // kernels.cu:175-178
                acc += (float)g.spin_no * blob_ptr<float>(B, L.deph_deg)[cnt_deph] / (float)A.n_spins_global;
                cnt_deph++;
            }
            if (mask_ev & EV_GRAD) { // kernels.cu:181-187
                const float Gx = __fmul_rn(gtx[cnt_grad], gscale), Gy = __fmul_rn(gty[cnt_grad], gscale), Gz = __fmul_rn(gtz[cnt_grad], gscale); // monte_carlo.cu:288-290
                const double X = (double)p0 * SC.unit_m[0], Y = (double)p1 * SC.unit_m[1], Z = (double)p2 * SC.unit_m[2];
                double gp = __fma_rn((double)Gz, Z, __fma_rn((double)Gx, X, __dmul_rn((double)Gy, Y)));
                gp = gp * 1e-3 * (double)A.timestep_us * 1e-6 * kGamma;
                if (MULTI) accg = (float)__fma_rn(gp, kRad2Deg, (double)accg);
                else acc = (float)__fma_rn(gp, kRad2Deg, (double)acc);
                cnt_grad++;
            }
            if (mask_ev & EV_RF) { // kernels.cu:190-199
                const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                const float rs = blob_ptr<float>(B, L.rf_s)[cur_rf], rc = blob_ptr<float>(B, L.rf_c)[cur_rf], rp = blob_ptr<float>(B, L.rf_ph)[cur_rf];
                relax_factors(dt_s);
                each_scale(false, [&](uint32_t, uint32_t, float *mm, const float gs, float) {
                    dephase_relax_pre(mm, fmaf(gs, accg, acc), relax, e1, e2);
                    float rr[3];
                    xrot_withphase(rs, rc, rp, mm, rr);
                    mm[0] = rr[0]; mm[1] = rr[1]; mm[2] = rr[2];
                });
                acc = 0.f; accg = 0.f;
                t_old = tp;
                cur_rf++;
            }
            if ((mask_ev & EV_ECHO) && scan + 1 == A.n_scans) { // kernels.cu:202-215
                const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                relax_factors(dt_s);
                each_scale(false, [&](uint32_t kk, uint32_t k_acc, float *mm, const float gs, float) {
                    dephase_relax_pre(mm, fmaf(gs, accg, acc), relax, e1, e2);
                    if (stage) *g.stage(A, cur_te, kk) = make_uint4(__float_as_uint(mm[0]), __float_as_uint(mm[1]), __float_as_uint(mm[2]), ts);
                    // (the scale is part of the key: lanes of a warp that run this loop at different iterations may arrive here together)
                    if (A.sums_fx) echo_sums_add(g.bsum, (k_acc * A.n_te + cur_te) * L.n_sub + ts, mm);
                });
                acc = 0.f; accg = 0.f;
                t_old = tp;
                cur_te++;
            }
            ev++;
        } else if (seg == SEG_RUN0) { // the plain steps before a run of gradient samples are done: now the run itself
            seg = SEG_RUN1;
            flags |= WF_GRUN;
            grad_first = cnt_grad;
            rem = (int)run_len;
            t_stop += run_len;
            break;
        } else if (seg == SEG_RUN1) {
            flags &= ~WF_GRUN;
            cnt_grad = grad_first + run_len;
            ev += run_len;
        } else {
            if (seg == SEG_TAIL) { // end of TR (kernels.cu:226-231)
                // RE-SYNCHRONISATION of multi-TR runs.  Every permeability rejection costs a walker one round, so the lanes of a warp reach their
                // events at different rounds and — over the ~1100 TRs of a bSSFP run — drift apart completely: the event code below would run once
                // per lane instead of once per warp.  TR number i therefore ends no earlier than round (i + 1) x A.tr_period, where the period
                // (engine.cu) is what a walker without rejections needs for a TR plus a slack of 2 + timepoints / 64 rounds: walkers that lose
                // fewer rounds than the slack per TR — nearly all, except behind walls at small FoV scales — stay on the common schedule, and a
                // walker that fell behind catches up by the slack of every TR.  Deterministic per walker (absolute round numbers).
                if (scan + 1 < A.n_scans && r_next < (scan + 1u) * A.tr_period) break; // called again at the next sync round (rem stays 0); also before a re-binning pause
                const float dt_s = (float)((double)((n_tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                if (scan + 1 < A.n_scans) { // (what the LAST scan leaves behind is read by nobody: kernels.cu:226-231 has no output after it)
                    relax_factors(dt_s);
                    each_scale(false, [&](uint32_t, uint32_t, float *mm, const float gs, float) { dephase_relax_pre(mm, fmaf(gs, accg, acc), relax, e1, e2); });
                }
                scan++;
                if (scan >= A.scan_end) { finished = true; break; }
            }
            { // start of a TR: phase cycling + first RF (kernels.cu:110-126)
                each_scale(seg == SEG_START, [&](uint32_t, uint32_t, float *mm, float, const float lin_pc) {
                    float ph = (float)((double)(A.rf_ph0 + (float)scan * lin_pc) + (double)(scan * (scan + 1u)) / 2.0 * (double)A.quad_pc);
                    // the reference wraps by repeated subtraction (kernels.cu:112-113: hundreds of iterations late in a bSSFP run): whole turns in closed form first
                    if (ph > 720.f) ph = (float)((double)ph - 360.0 * floor(((double)ph - 360.0) / 360.0));
                    if (ph < -360.f) ph = (float)((double)ph + 360.0 * floor(-(double)ph / 360.0));
                    while (ph > 360.0) ph = (float)(ph - 360.0);
                    while (ph < 0) ph = (float)(ph + 360.0);
                    float rr[3];
                    xrot_withphase(A.s, A.c, ph, mm, rr);
                    mm[0] = rr[0]; mm[1] = rr[1]; mm[2] = rr[2];
                });
                t_stop = 0; t_old = 0; ev = 0;
                cur_rf = 1; cur_te = 0; cnt_deph = 0; cnt_grad = 0;
                acc = 0.f; accg = 0.f;
                if (STATS) flags |= WF_FRESH;
            }
        }
        // ---- next segment: the steps up to and including the next entry's timepoint, or the plain steps before a run of
        //      gradient-only samples at consecutive timepoints (a PGSE lobe: one sample per step), or the rest of the TR ----
        const uint32_t ev_time = ev < L.n_tl ? (uint32_t)tl_time[ev] : n_tp;
        uint32_t stop;
        if (ev >= L.n_tl || ev_time >= n_tp) { stop = n_tp; seg = SEG_TAIL; }
        else if (GRUNS && tl_run[ev] >= 2u) { stop = ev_time; seg = SEG_RUN0; run_len = min(tl_run[ev], n_tp - ev_time); }
        else { stop = ev_time + 1u; seg = SEG_EVENT; }
        rem = (int)(stop - t_stop);
        t_stop = stop;
        if (rem > 0) break;
    }
    if (finished) { // this launch's scans are complete (or the walker was abandoned)
        flags |= WF_DONE;
        rem = 0;
        if (!MULTI && A.scan_end < A.n_scans) { // pause at a TR boundary: the round index is the whole RNG state (MULTI runs are never paused, engine.cu)
            A.state_a[g.st_idx(A)] = make_uint4(p0, p1, p2, r_next);
            A.state_b[g.st_idx(A)] = make_uint4(__float_as_uint(m[0]), __float_as_uint(m[1]), __float_as_uint(m[2]), ts | ((lost ? 1u : 0u) << 8));
            A.state_vox[g.st_idx(A)] = ((p0 >> SC.fb) * A.ny + (p1 >> SC.fb)) * A.nz + (p2 >> SC.fb);
        } else if (stage) {
            // echoes that never fired (beyond the TR, or after the walker was abandoned) read 0, like the reference's zero-initialised
            // outputs (monte_carlo.cu:256,259-260); the final position is the last committed one (kernels.cu:220-221)
            const uint4 pos = make_uint4(__float_as_uint((float)((double)p0 * SC.unit_m[0])), __float_as_uint((float)((double)p1 * SC.unit_m[1])),
                                         __float_as_uint((float)((double)p2 * SC.unit_m[2])), lost ? 1u : 0u);
            const uint32_t k0 = MULTI ? 0u : g.k, k1 = MULTI ? A.n_multi : g.k + 1u;
            for (uint32_t kk = k0; kk < k1; kk++) {
                for (uint32_t e = cur_te; e < A.n_te; e++) *g.stage(A, e, kk) = make_uint4(0u, 0u, 0u, 0u);
                if (!RECORD) *g.stage(A, A.n_te, kk) = pos;
            }
        }
    } else {
        if (!MULTI) { es[ES_M0 * nthr] = __float_as_uint(m[0]); es[ES_M1 * nthr] = __float_as_uint(m[1]); es[ES_M2 * nthr] = __float_as_uint(m[2]); }
        es[ES_SCAN * nthr] = scan; es[ES_EV * nthr] = ev; es[ES_SEG * nthr] = seg; es[ES_TSTOP * nthr] = t_stop; es[ES_TOLD * nthr] = t_old;
        es[ES_RF * nthr] = cur_rf; es[ES_TE * nthr] = cur_te; es[ES_DEPH * nthr] = cnt_deph; es[ES_GFIRST * nthr] = grad_first; es[ES_RUNLEN * nthr] = run_len;
    }
    AdvOutMulti o;
    o.acc = acc; o.rem = rem; o.cnt_grad = cnt_grad; o.flags = flags; o.accg = accg;
    return o;
}

// GRUNS: the sequence holds runs of gradient samples (taken inside the round); sequences without them get a kernel without that code.
// SHARED: block = 32 spins x A.group scales sharing the spins' normals through shared memory; else block = kBlock spins of one scale.
// MULTI (PRIVATE geometry, one "scale" per launch): one walker per spin for ALL A.n_multi gradient / phase-cycling scales (see advance_walker).
template <bool STATS, bool RECORD, int VOX, bool GRUNS, bool SHARED, bool MULTI = false>
__global__ void __launch_bounds__(SHARED ? SWK_FAST_SHARED_MAXT : kBlock, SHARED ? SWK_FAST_SHARED_MINB : (MULTI ? SWK_FAST_MULTI_MINB : SWK_FAST_MIN_BLOCKS))
walk_fast_kernel(const __grid_constant__ WalkArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const BlobLayout &L = A.L;
    const uint32_t lane = threadIdx.x & 31u;

    // ---- shared memory: [sequence tables] [block sums, int64] [scale constants of this block's scales] [event state] [normals, SHARED];
    //      which (spin, scale): see Geo ----
    const Geo<SHARED, MULTI> g(A, smem);
    const uint32_t nthr = g.nthr;
    const uint8_t *B = g.B;
    if (A.blob_in_smem) {
        const uint32_t nw = L.bytes / 4;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.blob);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem);
        for (uint32_t i = threadIdx.x; i < nw; i += nthr) dst[i] = __ldg(src + i);
    }
    static_assert(!(MULTI && (SHARED || RECORD)), "MULTI: PRIVATE geometry, no trajectory recording");
    const uint32_t n_grp = SHARED ? A.group : 1u; // scales walked by this block
    const uint32_t n_acc = MULTI ? A.n_multi : n_grp; // scales it accumulates ensemble sums for
    const uint32_t n_bsum = g.n_bsum;             // sum entries per scale
    long long *bsum = g.bsum;
    uint32_t *es = g.es;   // field f of this thread: es[f * nthr]
    float4 *nbuf = g.nbuf; // [2][kBatch][32], SHARED only
    const uint32_t k_first = g.k_first, kc = g.k, j = g.j, jl = g.jl, spin_no = g.spin_no;
    for (uint32_t i = threadIdx.x; i < n_bsum * n_acc; i += nthr) bsum[i] = 0;
    {
        const uint32_t wps = A.scale_stride / 4u; // words per scale record
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.scale_tab) + (size_t)k_first * wps;
        for (uint32_t i = threadIdx.x; i < wps * n_grp; i += nthr)
            reinterpret_cast<uint32_t *>(g.sct)[i] = (k_first + i / wps) < A.k_hi ? __ldg(src + i) : 0u;
    }
    __syncthreads();
    if (MULTI && GRUNS && A.g4_smem) { // gradient samples, pre-multiplied with the degrees of phase per (mT/m x fixed-point unit) of each axis: one LDS.128 per sample
        const ScaleConst &S0 = g.sc(A);
        const float *tx = blob_ptr<float>(B, L.gx), *ty = blob_ptr<float>(B, L.gy), *tz = blob_ptr<float>(B, L.gz);
        for (uint32_t i = threadIdx.x; i < L.n_grad; i += nthr) nbuf[i] = make_float4(tx[i] * S0.umk[0], ty[i] * S0.umk[1], tz[i] * S0.umk[2], 0.f);
        __syncthreads();
    }

    const bool valid = g.valid;
    const ScaleConst &SC = g.sc(A);
    const float *sgt = reinterpret_cast<const float *>(reinterpret_cast<const uint8_t *>(&SC) + SC.sgt_off);
    const uint32_t fb = SC.fb;

    const float *gtx = blob_ptr<float>(B, L.gx), *gty = blob_ptr<float>(B, L.gy), *gtz = blob_ptr<float>(B, L.gz);
    const float *tpXY = blob_ptr<float>(B, L.pXY);
    const uint32_t n0 = A.nx, n1 = A.ny, n2 = A.nz;
    const size_t st_idx = g.st_idx(A);
    (void)j;

    // ---- walker state ----
    uint32_t p0 = 0, p1 = 0, p2 = 0; // fixed-point position
    uint32_t r_first = 0;            // first round of this launch (a multiple of kSync)
    bool lost = false, lost_before = false;
    uint32_t ts_saved = 0;
    {
        float m[3] = {0.f, 0.f, 1.f};
        if (valid) {
            if (A.scan_first == 0) {
                uint32_t pp[3];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const float x0 = __ldg(A.xyz0 + 3 * (size_t)jl + i);
                    if (A.m0) m[i] = __ldg(A.m0 + 3 * (size_t)jl + i);
                    pp[i] = (uint32_t)fmin(fmax((double)x0 * SC.pos_k[i], 0.), SC.pos_hi[i]);
                }
                p0 = pp[0]; p1 = pp[1]; p2 = pp[2];
            } else { // resuming after a re-binning pause (engine.cu run_impl): position, round, magnetisation, substrate come back from the state arrays
                const uint4 sa = A.state_a[st_idx], sb = A.state_b[st_idx];
                p0 = sa.x; p1 = sa.y; p2 = sa.z; r_first = sa.w;
                m[0] = __uint_as_float(sb.x); m[1] = __uint_as_float(sb.y); m[2] = __uint_as_float(sb.z);
                ts_saved = sb.w & 0xffu;
                lost = lost_before = (sb.w & 0x100u) != 0u; // abandoned in an earlier launch (already counted there)
            }
        }
        if (MULTI) { // every scale starts from the same magnetisation: slot 0 carries it to the start of the first TR (advance_walker_multi, from_first)
            if (valid) A.mstate[j - A.m_first] = make_uint4(__float_as_uint(m[0]), __float_as_uint(m[1]), __float_as_uint(m[2]), 0u);
        } else {
            es[ES_M0 * nthr] = __float_as_uint(m[0]); es[ES_M1 * nthr] = __float_as_uint(m[1]); es[ES_M2 * nthr] = __float_as_uint(m[2]);
        }
        es[ES_SCAN * nthr] = A.scan_first; es[ES_SEG * nthr] = SEG_START;
        es[ES_EV * nthr] = 0u; es[ES_TSTOP * nthr] = 0u; es[ES_TOLD * nthr] = 0u; es[ES_RF * nthr] = 1u; es[ES_TE * nthr] = 0u;
        es[ES_DEPH * nthr] = 0u; es[ES_GFIRST * nthr] = 0u; es[ES_RUNLEN * nthr] = 0u;
    }
    // current voxel: table index, substrate, field.  The reference loads field / T1 / T2 at the first accepted step
    // (kernels.cu:91,150-170); holding the field of the CURRENT voxel from the start is equivalent: a first step that stays in
    // the voxel reads this very value.
    // index of a voxel in the table that is walked: the [nx][ny] slab, the row-major volume, or (packed words, A.brick) the volume cut into bricks
    // of 2 x 2 x 4 voxels = one 64-byte fetch unit each, so that a step of a fraction of a voxel along ANY axis mostly stays inside the line it has
    const uint32_t bry = A.brick ? (n1 + 1u) >> 1 : 0u, brz = A.brick ? (n2 + 3u) >> 2 : 0u;
    auto table_index = [&](const uint32_t v0, const uint32_t v1, const uint32_t v2) -> uint32_t {
        if (VOX == VOX_SLAB) return v0 * n1 + v1;
        if (VOX == VOX_PACKED && A.brick) return ((((v0 >> 1) * bry + (v1 >> 1)) * brz + (v2 >> 2)) << 4) | ((v0 & 1u) << 3) | ((v1 & 1u) << 2) | (v2 & 3u);
        return (v0 * n1 + v1) * n2 + v2;
    };
    uint32_t idx_cur = table_index(p0 >> fb, p1 >> fb, p2 >> fb);
    uint32_t ind3_cur = ((p0 >> fb) * n1 + (p1 >> fb)) * n2 + (p2 >> fb); // STATS only
    uint32_t wcur = 0;   // PACKED / SLAB: the packed word; MASK / SPLIT: the substrate id
    float fcur = 0.f;    // SPLIT: field of the current voxel (Tesla)
    if (valid) {
        if (VOX == VOX_PACKED || VOX == VOX_SLAB) wcur = __ldg(A.packed + idx_cur);
        else {
            wcur = __ldg(A.mask + idx_cur);
            if (VOX == VOX_SPLIT) fcur = __ldg(A.fieldmap + idx_cur);
        }
        if (A.scan_first > 0) { // the substrate a walker is IN (it equals its voxel's today: a rejected step is never taken)
            if (VOX == VOX_PACKED || VOX == VOX_SLAB) wcur = (wcur & ~15u) | ts_saved; else wcur = ts_saved;
        }
    }
    auto ts_of = [](uint32_t w) -> uint32_t { return (VOX == VOX_PACKED || VOX == VOX_SLAB) ? (w & 15u) : w; };
    float sg0 = sgt[3 * ts_of(wcur)], sg1 = sgt[3 * ts_of(wcur) + 1], sg2 = sgt[3 * ts_of(wcur) + 2];

    const size_t out_row = (size_t)kc * A.n_local + jl;
    float *X1 = (RECORD && A.XYZ1) ? A.XYZ1 + out_row * A.trj * 3 : nullptr;       // trajectories go straight to the reference layout
    if (RECORD && X1 && valid && A.scan_first == 0) { // slot 0 starts as the (scaled) initial position (kernels.cu:96)
#pragma unroll
        for (int i = 0; i < 3; i++) X1[i] = __fmul_rn(__ldg(A.xyz0 + 3 * (size_t)jl + i), SC.fscale);
    }

    // ---- registers of the round loop ----
    float acc = 0.f;      // phase accrued since the last event (degrees)
    float accg = 0.f;     // MULTI: its gradient part, for UNSCALED gradients (phase of scale k = acc + gscale_k accg)
    int rem = 0;          // accepted steps still to take in the current segment
    uint32_t itr = 0;     // consecutive rejections (kernels.cu:155)
    uint32_t cnt_grad = 0;
    bool grun = false;    // the current segment is a run of gradient samples (one per accepted step)
    bool done = !valid;
    bool fresh = true;    // STATS only (ind_old = matrix_length+1 at a TR start, kernels.cu:123)
    uint32_t st_mask = 0, st_field = 0, st_rej = 0, st_steps = 0; // per thread and launch: < 2^32
    const uint32_t n_tp = A.n_tp;
    const uint32_t seed_lo = (uint32_t)A.seed;
    const uint32_t seed_hi_walk = ((uint32_t)(A.seed >> 32) & 0x3fffffffu) | (STREAM_WALK << 30);
    const uint32_t kOne = A.one_bits; // 0x3f800000, deliberately opaque to ptxas (see and_or)

    // ======================================= one attempt =======================================
    // one tentative step (kernels.cu:130-170) with the normals of round `r`; `perm_u` yields the permeability uniform of the round;
    // `overlap` is a hook for independent work between ISSUING the voxel gather and CONSUMING it (unused today, see the PRIVATE loop below).
    auto attempt = [&](const float a0, const float a1, const float a2, const uint32_t r, auto &&perm_u, auto &&overlap) {
        const bool act = rem > 0;
        uint32_t q0 = fx_step(p0, a0, sg0), q1 = fx_step(p1, a1, sg1), q2 = fx_step(p2, a2, sg2);
        uint32_t v0 = q0 >> fb, v1 = q1 >> fb, v2 = q2 >> fb;
        if (act & ((v0 >= n0) | (v1 >= n1) | (v2 >= n2))) { // FoV boundary (kernels.cu:133-136), rare
            if (v0 >= n0) { q0 = fov_boundary(q0, p0, n0, fb, A.cross_fov); v0 = q0 >> fb; }
            if (v1 >= n1) { q1 = fov_boundary(q1, p1, n1, fb, A.cross_fov); v1 = q1 >> fb; }
            if (v2 >= n2) { q2 = fov_boundary(q2, p2, n2, fb, A.cross_fov); v2 = q2 >> fb; }
        }
        const uint32_t idx = table_index(v0, v1, v2);
        uint32_t w = wcur;
        float fv = fcur;
        if (act & (idx != idx_cur)) { // the voxel changed: one gather
            if (VOX == VOX_PACKED || VOX == VOX_SLAB) w = ldg_voxel(A.packed + idx);
            else {
                w = __ldg(A.mask + idx);
                if (VOX == VOX_SPLIT) fv = __ldg(A.fieldmap + idx);
            }
        }
        overlap();
        bool chg = false;
        uint32_t ind3 = 0;
        if (STATS) {
            ind3 = (v0 * n1 + v1) * n2 + v2;
            chg = act & ((ind3 != ind3_cur) | fresh);
            st_mask += chg;
        }
        bool ok = act;
        if ((VOX == VOX_PACKED || VOX == VOX_SLAB) ? (((w ^ wcur) & 15u) != 0u) : (w != wcur)) { // kernels.cu:150-164 (only a lane that hopped gets here)
            // accept iff u < P_XY[from][to], u in [0,1) (kernels.cu:154).  P <= 0 always rejects and P >= 1 always accepts.
            const float pxy = tpXY[ts_of(wcur) * L.n_sub + ts_of(w)];
            bool reject = pxy <= 0.f;
            if (pxy > 0.f && pxy < 1.f) reject = perm_u(r) >= pxy;
            if (reject) {
                ok = false;
                if (STATS) st_rej++;
                if (itr++ > A.max_iter) { lost = true; rem = 0; } // kernels.cu:155-159
            } else {
                sg0 = sgt[3 * ts_of(w)]; sg1 = sgt[3 * ts_of(w) + 1]; sg2 = sgt[3 * ts_of(w) + 2];
            }
        }
        if (ok) { // commit (kernels.cu:165-172, 218-223)
            p0 = q0; p1 = q1; p2 = q2;
            idx_cur = idx;
            wcur = w;
            fcur = fv;
            if (VOX == VOX_PACKED || VOX == VOX_SLAB) acc = fmaf(__uint_as_float(w), A.field_k, acc); // kernels.cu:171-172, monte_carlo.cu:244
            else if (VOX == VOX_SPLIT) acc = fmaf(fv, A.field_k, acc);
            itr = 0;
            rem--;
            if (STATS) { st_field += chg; ind3_cur = ind3; fresh = false; st_steps++; }
            if (GRUNS && grun) { // gradient sample of this timepoint, at the NEW position (kernels.cu:181-187); FP32 here, FP64 in the event path
                if (MULTI) {
                    if (A.g4_smem) {
                        const float4 g4 = nbuf[cnt_grad];
                        accg += g4.x * (float)p0 + g4.y * (float)p1 + g4.z * (float)p2;
                    } else accg += gtx[cnt_grad] * ((float)p0 * SC.umk[0]) + gty[cnt_grad] * ((float)p1 * SC.umk[1]) + gtz[cnt_grad] * ((float)p2 * SC.umk[2]);
                } else {
                    const float gs = SC.gscale;
                    const float gx = __fmul_rn(gtx[cnt_grad], gs), gy = __fmul_rn(gty[cnt_grad], gs), gz = __fmul_rn(gtz[cnt_grad], gs);
                    acc += gx * ((float)p0 * SC.umk[0]) + gy * ((float)p1 * SC.umk[1]) + gz * ((float)p2 * SC.umk[2]);
                }
                cnt_grad++;
            }
            if (RECORD && X1) { // kernels.cu:218-221 (diagnostic mode)
                float *slot = X1 + 3 * ((size_t)es[ES_SCAN * nthr] * n_tp + (es[ES_TSTOP * nthr] - 1u - (uint32_t)rem));
                slot[0] = (float)((double)p0 * SC.unit_m[0]); slot[1] = (float)((double)p1 * SC.unit_m[1]); slot[2] = (float)((double)p2 * SC.unit_m[2]);
            }
        }
    };

    // ======================================= events =======================================
    auto advance = [&](const uint32_t r_next) { // (see advance_walker)
        const uint32_t fl = (lost ? WF_LOST : 0u) | (grun ? WF_GRUN : 0u) | (fresh ? WF_FRESH : 0u);
        if constexpr (MULTI) {
            const AdvOutMulti o = advance_walker_multi<STATS, VOX, GRUNS>(&A, p0, p1, p2, wcur, acc, accg, cnt_grad, fl, r_next);
            acc = o.acc; rem = o.rem; cnt_grad = o.cnt_grad; accg = o.accg;
            done = (o.flags & WF_DONE) != 0u; grun = (o.flags & WF_GRUN) != 0u;
            if (STATS) fresh = (o.flags & WF_FRESH) != 0u;
        } else {
            const AdvOut o = advance_walker<STATS, RECORD, VOX, GRUNS, SHARED>(&A, p0, p1, p2, wcur, acc, cnt_grad, fl, r_next);
            acc = o.acc; rem = o.rem; cnt_grad = o.cnt_grad;
            done = (o.flags & WF_DONE) != 0u; grun = (o.flags & WF_GRUN) != 0u;
            if (STATS) fresh = (o.flags & WF_FRESH) != 0u;
        }
    };

    if (!done) advance(r_first); // start of the first TR of this launch

    if (SHARED) {
        auto generate = [&](float4 *buf, const uint32_t r0) {
            generate_normals(buf, r0, spin_no, seed_lo, seed_hi_walk, kOne, A.perm_draws ? A.perm_key : nullptr);
        };
        uint32_t r0 = 0, cur = 0;
        generate(nbuf, 0u);
        for (;;) {
            if (!__syncthreads_or(!done)) break; // also: batch `cur` is visible, and everybody has finished reading the other buffer
            generate(nbuf + (cur ^ 1u) * (kBatch * 32u), r0 + kBatch);
            if (!done) {
#pragma unroll 1
                for (uint32_t h = 0; h < kBatch; h += kSync) {
                    const float4 *nb = nbuf + cur * (kBatch * 32u) + h * 32u + lane;
#pragma unroll kUnroll
                    for (uint32_t rr = 0; rr < kSync; rr++) {
                        const float4 c = nb[rr * 32u];
                        attempt(c.x, c.y, c.z, r0 + h + rr, [&](uint32_t) { return c.w; }, [] {});
                    }
                    if (rem == 0) advance(r0 + h + kSync);
                    if (done) break;
                }
            }
            r0 += kBatch;
            cur ^= 1u;
        }
    } else {
        // (Overlapping the next block's Philox rounds / Box-Muller with the gathers of these two attempts, as the round-1 kernel did, was measured
        // again on this kernel: the ten extra live registers spill inside the loop and the walk gets 19 % slower; profiles/README.md.)
        uint32_t r = r_first;
        auto pu = [&](uint32_t rr) { return u01_open1(philox2x32_10(rr, spin_no, A.perm_key)); };
        while (!done) {
#pragma unroll 1
            for (uint32_t h = 0; h < kSync; h += 2u) {
                float a0, a1, a2, b0, b1, b2;
                normals6_fast(philox_fixed(r >> 1, seed_lo, spin_no, seed_hi_walk), kOne, a0, a1, a2, b0, b1, b2);
                attempt(a0, a1, a2, r, pu, [] {});
                attempt(b0, b1, b2, r + 1u, pu, [] {});
                r += 2u;
            }
            if (rem == 0) advance(r);
        }
    }

    // ---- flush block sums and counters ----
    __syncthreads();
    if (A.sums_fx) {
        for (uint32_t i = threadIdx.x; i < n_bsum * n_acc; i += nthr) {
            const uint32_t kk = (MULTI ? 0u : k_first) + i / n_bsum;
            const long long v = bsum[i];
            if (v != 0 && kk < (MULTI ? A.n_multi : A.k_hi)) atomicAdd(A.sums_fx + (size_t)kk * n_bsum + (i % n_bsum), (unsigned long long)v);
        }
    }
    if (A.counters) {
        if (STATS) {
            unsigned long long c0 = st_steps, c1 = st_mask, c2 = st_field, c3 = st_rej;
            if (MULTI) { c0 *= A.n_multi; c1 *= A.n_multi; c2 *= A.n_multi; c3 *= A.n_multi; } // the statistics of the n_multi identical walks this one stands for
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                c2 += __shfl_xor_sync(0xffffffffu, c2, o);
                c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            }
            if (lane == 0) {
                atomicAdd(A.counters + 0, c0);
                atomicAdd(A.counters + 1, c1);
                atomicAdd(A.counters + 2, c2);
                atomicAdd(A.counters + 3, c3);
            }
        }
        const unsigned lost_w = __popc(__ballot_sync(0xffffffffu, lost && !lost_before)) * (MULTI ? A.n_multi : 1u); // spins abandoned in THIS launch
        if (lane == 0 && lost_w) atomicAdd(A.counters + 4, (unsigned long long)lost_w);
    }
}

} // namespace swk
