// spinwalk_b200/csrc/phantom.cuh — phantom generator of SURVEY §8 row f3: host placement + CUDA voxel fill.
//
// What the reference does (src/phantom/*.cpp, one OpenMP loop nest per shape over that shape's bounding box, reading a
// 12 B/voxel coordinate grid and read-modify-writing the field map once per shape) is restated per VOXEL: one thread owns a
// voxel (spheres) or an (x,y) column (cylinders, whose mask and field do not depend on z), visits the shapes in acceptance
// order with the field value in a register, and writes each output byte exactly once.  Summation order per voxel is the
// reference's (shape 0, 1, 2, ...), every float/double operation is issued with an explicit round-to-nearest intrinsic so
// that nothing is contracted into an FMA, and expression types follow the reference (float geometry, the field term in
// double because of the M_PI literal, `+=` into a float) => bit-identical masks and field maps.
//
//   cylinders  cyl_slab_kernel      [res][res] slab of (mask, field) per column          compute, ~res^2 x n_cyl, tiny
//              slab_broadcast_bulk_kernel slab -> [res][res][res] volume: 4096-voxel chunks assembled in shared memory and written
//                                    with TMA bulk stores (cp.async.bulk shared -> global)  HBM-write bound: 5 B per voxel
//              slab_broadcast_kernel the same with plain 16 B stores (A/B reference, SWK_PHANTOM_BCAST=stg)
//              cyl_exact_kernel      only for columns where the reference's z residual (below) could flip a rounding
//   spheres    sphere_fill_kernel    8x8x32 tile per block, shapes filtered per tile     FP64-divide bound (two IEEE
//                                    (order-preserving ballot compaction into smem)      double divisions per voxel-shape pair)
//   two pools  two cudaMemsetAsync
//
// The z residual: the reference projects grid-point minus cylinder-point onto the axis and subtracts again
// (phantom_cylinder.cpp:250-252), which leaves perpendicular[2] = gz - fl(fl(gz - cz) + cz), a rounding residue of ~1e-6 µm
// that enters |perpendicular|.  It changes fl(distance2 + residue^2) only when distance2 is within a few ulps of nothing, so
// the slab kernel tests fl(distance2 + max_z residue^2) == distance2 per (column, cylinder) and flags the column otherwise.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/spinwalk_phantom.h"

namespace swk {
namespace phantom {

struct Shape { float x, y, z, r; };

// ---------------------------------------------------------------- host: grid + placement ----------------------------------------------------------------

// voxel centres (phantom_base.cpp:122-129): start/end/step in double from float fov / size_t resolution, stored as float
inline std::vector<float> voxel_centres(float fov, size_t res)
{
    std::vector<float> g(res);
    const double first = fov / res / 2.0;
    const double last = fov - fov / res / 2.0;
    const double pitch = (last - first) / (res - 1.0);
    for (size_t i = 0; i < res; i++) g[i] = static_cast<float>(first + i * pitch);
    return g;
}

// Uniform grid over the shape centres for the collision test.  The reference scans every placed shape for every candidate
// (O(N) per candidate, 44 k spheres for the C3 phantom); here only the shapes of nearby cells are visited — in their ORIGINAL
// index order, because with random radii the candidate shrinks step by step and the outcome depends on the visiting order.
// Why the others cannot matter: a placed shape (radius R <= rmax) interacts with a candidate (radius r <= rmax) only when
// dist <= R, dist <= r or dist < fl(R + r), i.e. dist <= 2 rmax (+ rounding); cells are 1.0001 rmax wide (or wider when that
// would give too many cells) and the query takes 2 cells either side, so a shape outside the query is > 2.0002 rmax away along an axis.
template <int DIM>
class CentreGrid {
public:
    CentreGrid(double lo, double hi, double rmax)
    {
        origin_ = lo;
        cell_ = std::max(rmax * 1.0001, (hi - lo) / (DIM == 2 ? 1024.0 : 192.0));
        if (!(cell_ > 0)) cell_ = 1.0;
        n_ = int(std::floor((hi - lo) / cell_)) + 1;
        cells_.resize(DIM == 2 ? size_t(n_) * n_ : size_t(n_) * n_ * n_);
    }
    void add(uint32_t index, const float *c) { cells_[flat(cell_of(c[0]), cell_of(c[1]), DIM == 3 ? cell_of(c[2]) : 0)].push_back(index); }
    // calls visit(index) for every shape within 2 cells of c (any order)
    template <class F>
    void for_nearby(const float *c, F visit) const
    {
        const int cx = cell_of(c[0]), cy = cell_of(c[1]), cz = DIM == 3 ? cell_of(c[2]) : 0;
        const int z0 = DIM == 3 ? std::max(0, cz - 2) : 0, z1 = DIM == 3 ? std::min(n_ - 1, cz + 2) : 0;
        for (int x = std::max(0, cx - 2); x <= std::min(n_ - 1, cx + 2); x++)
            for (int y = std::max(0, cy - 2); y <= std::min(n_ - 1, cy + 2); y++)
                for (int z = z0; z <= z1; z++)
                    for (const uint32_t i : cells_[flat(x, y, z)]) visit(i);
    }

private:
    int cell_of(float x) const { return std::min(n_ - 1, std::max(0, int(std::floor((double(x) - origin_) / cell_)))); }
    size_t flat(int x, int y, int z) const { return DIM == 2 ? size_t(x) * n_ + y : (size_t(x) * n_ + y) * n_ + z; }
    double origin_, cell_;
    int n_;
    std::vector<std::vector<uint32_t>> cells_;
};

// Does candidate c (radius may shrink when radii are random) collide with the shapes placed so far?
// DIM = 2: distance in the xy plane (parallel cylinders, phantom_cylinder.cpp:22-56); DIM = 3: spheres (phantom_sphere.cpp:23-55).
// The reference visits ALL placed shapes in index order (serially; its OpenMP build races on `radius`, the result then depends
// on thread timing).  The same outcome with far fewer visits:
//   1. a shape can only take part while dist <= R, dist <= r or dist < fl(R + r) for the candidate's current radius r, and r
//      never grows beyond its initial value r0 (up to a rounding of fl(dist - R)), so shapes with dist > R + r0 + margin are
//      skipped — `dist` being the reference's own float expression;
//   2. the survivors (a handful) are sorted by index and taken through the reference's sequential rule.
template <int DIM>
inline bool collides(const std::vector<Shape> &placed, const CentreGrid<DIM> &grid, std::vector<std::pair<uint32_t, float>> &scratch, const float *c,
                     float &radius, bool random_radius)
{
    scratch.clear();
    const float reach = radius * 1.001f + 1e-4f;
    bool inside_one = false; // centre inside a placed shape: rejected whatever the order
    grid.for_nearby(c, [&](uint32_t i) {
        const Shape &s = placed[i];
        const float d0 = c[0] - s.x, d1 = c[1] - s.y;
        float dist;
        if (DIM == 2) dist = std::sqrt(d0 * d0 + d1 * d1);
        else {
            const float d2 = c[2] - s.z;
            dist = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        }
        if (dist <= s.r) inside_one = true;
        else if (dist <= s.r + reach) scratch.emplace_back(i, dist);
    });
    if (inside_one) return true;
    std::sort(scratch.begin(), scratch.end());
    for (const auto &[i, dist] : scratch) {
        const float R = placed[i].r;
        if (dist <= radius) return true;
        if (dist < R + radius) {
            if (!random_radius) return true;
            radius = dist - R;
        }
    }
    return false;
}

// volume a cylinder adds to the FoV (phantom_cylinder.cpp:133-181): analytic when it lies inside with a 1.5 µm margin,
// else counted on the voxel grid; < 0 when it misses the FoV.
inline float cylinder_volume(const std::vector<float> &g, float fov, size_t res, const float *c, float rad)
{
    for (int i = 0; i < 2; i++)
        if (c[i] + rad < 0 || c[i] - rad > fov) return -1.f;
    bool cut = false;
    for (int i = 0; i < 2; i++)
        cut = cut || c[i] < rad - 1.5 || c[i] > fov - rad + 1.5;
    if (!cut) return static_cast<float>(M_PI * rad * rad * fov);
    const float h = fov / res, rad2 = rad * rad;
    const int32_t vx = int32_t(c[0] / h), vy = int32_t(c[1] / h), rv = int32_t(std::ceil(rad / fov * res) + 1);
    const int32_t x0 = std::max(0, vx - rv), x1 = std::min(int32_t(res), vx + rv + 2);
    const int32_t y0 = std::max(0, vy - rv), y1 = std::min(int32_t(res), vy + rv + 2);
    int32_t inside = 0;
    for (int32_t py = y0; py < y1; py++)
        for (int32_t px = x0; px < x1; px++) {
            const float a = g[px] - c[0], b = g[py] - c[1];
            if (a * a + b * b <= rad2) inside++;
        }
    inside *= int32_t(res); // every z slice counts the same voxels
    return inside * h * h * h;
}

// The reference's placement loops have no exit when the target cannot be met (e.g. the remaining volume is smaller than any
// candidate the grid can count): `spinwalk phantom` then spins forever.  This engine gives up after this many consecutive
// rejected candidates and reports an error instead — the only deliberate deviation in the placement.
constexpr int SWK_PLACE_STALLED = -1;
inline uint64_t max_consecutive_rejections()
{
    const char *v = getenv("SWK_PHANTOM_MAX_REJECTIONS"); // tests shorten the wait
    const uint64_t n = v ? strtoull(v, nullptr, 10) : 0;
    return n ? n : 50ull * 1000 * 1000;
}

inline uint64_t resolve_seed(int32_t seed) { return seed >= 0 ? uint64_t(seed) : uint64_t(std::random_device{}()); }

// phantom_cylinder.cpp:85-130
inline int place_cylinders(const swk_phantom_spec &sp, const std::vector<float> &g, std::vector<Shape> &out)
{
    const float fov = sp.fov_um, vf = sp.volume_fraction;
    if (2 * sp.radius_um >= fov) return SWK_ERR_INVALID;
    const bool random_radius = sp.radius_um < 0;
    const float rmax = std::fabs(sp.radius_um);
    std::mt19937 gen(resolve_seed(sp.seed));
    std::uniform_real_distribution<float> u01(0.f, 1.f);
    float filled = 0, whole = fov * fov * fov;
    out.clear();
    CentreGrid<2> grid(-double(rmax), double(fov) + rmax, rmax); // centres lie in [-r, fov + r]
    std::vector<std::pair<uint32_t, float>> near;
    const uint64_t kMaxConsecutiveRejections = max_consecutive_rejections();
    uint64_t rejected = 0;
    for (int32_t percent = 0; percent < 100;) {
        if (rejected++ > kMaxConsecutiveRejections) return SWK_PLACE_STALLED;
        float rad = random_radius ? u01(gen) * rmax : rmax;
        float c[3];
        for (float &v : c) v = u01(gen) * (fov + 2 * rad) - rad;
        if (collides<2>(out, grid, near, c, rad, random_radius)) continue;
        const float vol = cylinder_volume(g, fov, sp.resolution, c, rad);
        if (100 * (vol + filled) / whole > 1.02 * vf || vol < 0) continue;
        rejected = 0;
        filled += vol;
        percent = int32_t(100 * (100. * filled / whole / vf));
        grid.add(uint32_t(out.size()), c);
        out.push_back({c[0], c[1], c[2], rad});
    }
    return SWK_OK;
}

// phantom_sphere.cpp:79-119
inline int place_spheres(const swk_phantom_spec &sp, std::vector<Shape> &out)
{
    const float fov = sp.fov_um, vf = sp.volume_fraction;
    if (2 * sp.radius_um >= fov) return SWK_ERR_INVALID;
    const bool random_radius = sp.radius_um < 0;
    const float rmax = std::fabs(sp.radius_um);
    std::minstd_rand gen(resolve_seed(sp.seed));
    std::uniform_real_distribution<float> u01(0.f, 1.f);
    float filled = 0, whole = fov * fov * fov;
    out.clear();
    CentreGrid<3> grid(0.0, double(fov), rmax); // centres lie in [0, fov)
    std::vector<std::pair<uint32_t, float>> near;
    const uint64_t kMaxConsecutiveRejections = max_consecutive_rejections();
    uint64_t rejected = 0;
    for (int32_t percent = 0; percent < 100;) {
        if (rejected++ > kMaxConsecutiveRejections) return SWK_PLACE_STALLED;
        float rad = random_radius ? u01(gen) * rmax : rmax;
        float c[3];
        for (float &v : c) v = u01(gen) * fov;
        if (collides<3>(out, grid, near, c, rad, random_radius)) continue;
        rejected = 0;
        filled += 4 * M_PI / 3 * rad * rad * rad;
        grid.add(uint32_t(out.size()), c);
        out.push_back({c[0], c[1], c[2], rad});
        percent = int32_t(0.95 * 100 * (100. * filled / whole / vf)); // 0.95: spheres cut by the FoV faces
    }
    return SWK_OK;
}

// ---------------------------------------------------------------- device data ----------------------------------------------------------------

struct CylDev {
    float cx, cy, cz, r2;
    float zres2_max;       // max over z of fl(residual^2), residual = gz - fl(fl(gz - cz) + cz)
    int32_t x0, x1, y0, y1; // bounding box of the reference's loop nest (phantom_cylinder.cpp:224-237)
};

struct CylConst {
    double k_out;    // 2*M_PI*(1-Y)*dChi                         (phantom_cylinder.cpp:259)
    double k_in;     // k_out * (cos^2(theta) - 1/3)              (phantom_cylinder.cpp:261)
    float b0x, b0y;  // B0 rotated about y by the orientation, projected on the xy plane, normalised (:76-80,200-201)
    float sin2;      // 1 - cos^2(theta)                          (:205)
};

struct SphDev {
    float cx, cy, cz, r2;
    double amp;            // 4*M_PI*(1-Y)*dChi * r^2 * r   (phantom_sphere.cpp:182)
    int32_t lo[3], hi[3];  // bounding box (phantom_sphere.cpp:147-165)
};

// ---------------------------------------------------------------- kernels ----------------------------------------------------------------

__device__ __forceinline__ float sq_rn(float a) { return __fmul_rn(a, a); }

// One cylinder's contribution to a voxel outside it.  d2 = distance^2 in the xy plane, zres2 = fl(perpendicular_z^2).
__device__ __forceinline__ double cyl_term(const CylConst &k, float p0, float p1, float d2, float zres2, float r2)
{
    const float nrm = __fsqrt_rn(__fadd_rn(d2, zres2));
    const float cphi = __fdiv_rn(__fadd_rn(__fmul_rn(p0, k.b0x), __fmul_rn(p1, k.b0y)), nrm);
    const float c2phi = __fsub_rn(__fmul_rn(__fmul_rn(2.f, cphi), cphi), 1.f);
    return __dmul_rn(__dmul_rn(__dmul_rn(k.k_out, (double)__fdiv_rn(r2, d2)), (double)c2phi), (double)k.sin2);
}

// [res][res] slab: thread = (x, y) column, y fastest; a block owns 256 consecutive columns and keeps only the cylinders whose
// box meets them (order-preserving ballot compaction, as in sphere_fill_kernel).  counters[0] += number of masked columns.
template <bool CALC>
__global__ void __launch_bounds__(256) cyl_slab_kernel(const float *__restrict__ g, const CylDev *__restrict__ cyl, uint32_t n_cyl, uint32_t res,
                                                       CylConst k, uint8_t *__restrict__ mask2, float *__restrict__ field2,
                                                       uint32_t *__restrict__ exact_list, unsigned int *__restrict__ counters /*[0]=ones [1]=exact*/)
{
    __shared__ CylDev s_cyl[256];
    __shared__ uint32_t s_warp_hits[8];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_cols = res * res;
    const uint32_t first = blockIdx.x * 256, last = min(first + 255u, n_cols - 1);
    const uint32_t col = first + threadIdx.x;
    const bool live = col < n_cols;
    const int32_t px = live ? int32_t(col / res) : 0, py = live ? int32_t(col % res) : 0;
    // the block's columns: x rows [bx0, bx1]; y range only when they share one row
    const int32_t bx0 = int32_t(first / res), bx1 = int32_t(last / res);
    const int32_t by0 = bx0 == bx1 ? int32_t(first % res) : 0, by1 = bx0 == bx1 ? int32_t(last % res) : int32_t(res) - 1;
    const float gx = g[px], gy = g[py];
    float f = 0.f;
    bool in_shape = false, exact = false;
    for (uint32_t base = 0; base < n_cyl; base += 256) {
        const uint32_t i = base + threadIdx.x;
        CylDev mine;
        bool hit = false;
        if (i < n_cyl) {
            mine = cyl[i];
            hit = mine.x0 <= bx1 && mine.x1 > bx0 && mine.y0 <= by1 && mine.y1 > by0;
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, hit);
        __syncthreads(); // previous batch fully consumed
        if (lane == 0) s_warp_hits[warp] = __popc(vote);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < 8; w++) {
            const uint32_t c = s_warp_hits[w];
            before += w < warp ? c : 0;
            total += c;
        }
        if (hit) s_cyl[before + __popc(vote & ((1u << lane) - 1u))] = mine;
        __syncthreads();
        for (uint32_t j = 0; j < total; j++) {
            const CylDev c = s_cyl[j];
            if (px < c.x0 || px >= c.x1 || py < c.y0 || py >= c.y1) continue;
            const float p0 = __fsub_rn(gx, c.cx), p1 = __fsub_rn(gy, c.cy);
            const float d2 = __fadd_rn(sq_rn(p0), sq_rn(p1));
            if (d2 <= c.r2) in_shape = true;
            if (CALC) {
                if (d2 > c.r2) {
                    exact = exact || __fadd_rn(d2, c.zres2_max) != d2;
                    f = __double2float_rn(__dadd_rn((double)f, cyl_term(k, p0, p1, d2, 0.f, c.r2)));
                } else
                    f = __double2float_rn(__dadd_rn((double)f, k.k_in));
            }
        }
    }
    if (live) {
        mask2[col] = in_shape ? 1 : 0;
        if (CALC) field2[col] = f;
        if (exact) exact_list[atomicAdd(&counters[1], 1u)] = col;
    }
    const unsigned int n1 = __syncthreads_count(live && in_shape);
    if (threadIdx.x == 0 && n1) atomicAdd(&counters[0], n1);
}

// slab[col] -> volume[col*res + z] for all z.  A block streams 4096 consecutive voxels per iteration: one 16 B mask store
// and four coalesced 16 B field stores per thread.  One 64-bit division per chunk; a store that lies inside one column (nearly
// all of them) is a splat of one slab value, the others walk their 4 / 16 voxels.  V % 4096 voxels at the end go one by one.
template <bool CALC>
__global__ void __launch_bounds__(256) slab_broadcast_kernel(const uint8_t *__restrict__ mask2, const float *__restrict__ field2, uint32_t res, uint64_t V,
                                                             uint8_t *__restrict__ mask, float *__restrict__ field)
{
    const uint64_t n_chunks = V / 4096;
    const uint32_t last_col = res * res - 1;
    const uint32_t step_col = 1024u / res, step_z = 1024u % res; // a field store of the next round lies 1024 voxels further
    for (uint64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const uint64_t v0 = ch * 4096;
        const uint64_t col0 = v0 / res;
        const uint32_t z0 = uint32_t(v0 - col0 * res);
        { // mask: voxels [v0 + 16 t, +16)
            const uint32_t off = z0 + 16u * threadIdx.x;
            uint32_t col = uint32_t(col0) + off / res, z = off % res;
            uint32_t mv = mask2[min(col, last_col)];
            uint4 w;
            if (z + 16u <= res) w.x = w.y = w.z = w.w = mv * 0x01010101u;
            else {
                uint32_t b[4] = {0, 0, 0, 0};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    b[i >> 2] |= mv << (8 * (i & 3));
                    if (++z == res) { z = 0; col++; mv = mask2[min(col, last_col)]; }
                }
                w = make_uint4(b[0], b[1], b[2], b[3]);
            }
            *reinterpret_cast<uint4 *>(mask + v0 + 16u * threadIdx.x) = w;
        }
        if (CALC) {
            const uint32_t off = z0 + 4u * threadIdx.x;
            uint32_t col = uint32_t(col0) + off / res, z = off % res;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float fv = field2[min(col, last_col)];
                float4 o;
                if (z + 4u <= res) o = make_float4(fv, fv, fv, fv);
                else {
                    float t[4];
                    uint32_t c2 = col, z2 = z;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        t[i] = fv;
                        if (++z2 == res) { z2 = 0; c2++; fv = field2[min(c2, last_col)]; }
                    }
                    o = make_float4(t[0], t[1], t[2], t[3]);
                }
                *reinterpret_cast<float4 *>(field + v0 + 4u * (q * 256 + threadIdx.x)) = o;
                col += step_col;
                z += step_z;
                if (z >= res) { z -= res; col++; }
            }
        }
    }
    // tail
    for (uint64_t v = n_chunks * 4096 + blockIdx.x * 256ull + threadIdx.x; v < V; v += uint64_t(gridDim.x) * 256) {
        const uint64_t col = v / res;
        mask[v] = mask2[col];
        if (CALC) field[v] = field2[col];
    }
}

// The same broadcast with TMA bulk stores: a block assembles a 4096-voxel chunk (4 KB of mask, 16 KB of field) in shared memory and
// one thread hands it to the copy engine (cp.async.bulk shared::cta -> global, SASS UBLKCP); two stages, so that the next chunk is
// assembled while the previous one drains.  The LSU then only sees shared-memory stores, and HBM sees whole 4 KB / 16 KB bursts.
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

template <bool CALC>
__global__ void __launch_bounds__(256) slab_broadcast_bulk_kernel(const uint8_t *__restrict__ mask2, const float *__restrict__ field2, uint32_t res, uint64_t V,
                                                                  uint8_t *__restrict__ mask, float *__restrict__ field)
{
    extern __shared__ __align__(128) uint8_t stage_mem[]; // 2 x (16384 B field + 4096 B mask)
    const uint64_t n_chunks = V / 4096;
    const uint32_t last_col = res * res - 1;
    const uint32_t step_col = 1024u / res, step_z = 1024u % res;
    uint32_t it = 0;
    for (uint64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, it++) {
        uint8_t *sm = stage_mem + (it & 1u) * 20480u;
        float4 *sf = reinterpret_cast<float4 *>(sm);
        uint4 *smk = reinterpret_cast<uint4 *>(sm + 16384);
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); // the stores that last read this stage are done with it
        __syncthreads();
        const uint64_t v0 = ch * 4096;
        const uint64_t col0 = v0 / res;
        const uint32_t z0 = uint32_t(v0 - col0 * res);
        {
            const uint32_t off = z0 + 16u * threadIdx.x;
            uint32_t col = uint32_t(col0) + off / res, z = off % res;
            uint32_t mv = mask2[min(col, last_col)];
            uint4 w;
            if (z + 16u <= res) w.x = w.y = w.z = w.w = mv * 0x01010101u;
            else {
                uint32_t b[4] = {0, 0, 0, 0};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    b[i >> 2] |= mv << (8 * (i & 3));
                    if (++z == res) { z = 0; col++; mv = mask2[min(col, last_col)]; }
                }
                w = make_uint4(b[0], b[1], b[2], b[3]);
            }
            smk[threadIdx.x] = w;
        }
        if (CALC) {
            const uint32_t off = z0 + 4u * threadIdx.x;
            uint32_t col = uint32_t(col0) + off / res, z = off % res;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float fv = field2[min(col, last_col)];
                float4 o;
                if (z + 4u <= res) o = make_float4(fv, fv, fv, fv);
                else {
                    float t[4];
                    uint32_t c2 = col, z2 = z;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        t[i] = fv;
                        if (++z2 == res) { z2 = 0; c2++; fv = field2[min(c2, last_col)]; }
                    }
                    o = make_float4(t[0], t[1], t[2], t[3]);
                }
                sf[q * 256 + threadIdx.x] = o;
                col += step_col;
                z += step_z;
                if (z >= res) { z -= res; col++; }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes above -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store(mask + v0, smk, 4096u);
            if (CALC) bulk_store(field + v0, sf, 16384u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // shared memory must outlive the copies
    __syncthreads();
    for (uint64_t v = n_chunks * 4096 + blockIdx.x * 256ull + threadIdx.x; v < V; v += uint64_t(gridDim.x) * 256) {
        const uint64_t col = v / res;
        mask[v] = mask2[col];
        if (CALC) field[v] = field2[col];
    }
}

// Voxel-by-voxel evaluation of the flagged columns with the z residual in place.  blockIdx.x = flagged column, threads over z.
__global__ void __launch_bounds__(256) cyl_exact_kernel(const float *__restrict__ g, const CylDev *__restrict__ cyl, uint32_t n_cyl, uint32_t res, CylConst k,
                                                        const uint32_t *__restrict__ exact_list, float *__restrict__ field)
{
    const uint32_t col = exact_list[blockIdx.x];
    const int32_t px = int32_t(col / res), py = int32_t(col % res);
    const float gx = g[px], gy = g[py];
    for (uint32_t pz = threadIdx.x; pz < res; pz += 256) {
        const float gz = g[pz];
        float f = 0.f;
        for (uint32_t j = 0; j < n_cyl; j++) {
            const CylDev c = cyl[j];
            if (px < c.x0 || px >= c.x1 || py < c.y0 || py >= c.y1) continue;
            const float p0 = __fsub_rn(gx, c.cx), p1 = __fsub_rn(gy, c.cy);
            const float d2 = __fadd_rn(sq_rn(p0), sq_rn(p1));
            if (d2 > c.r2) {
                const float zr = __fsub_rn(gz, __fadd_rn(__fsub_rn(gz, c.cz), c.cz));
                f = __double2float_rn(__dadd_rn((double)f, cyl_term(k, p0, p1, d2, sq_rn(zr), c.r2)));
            } else
                f = __double2float_rn(__dadd_rn((double)f, k.k_in));
        }
        field[uint64_t(col) * res + pz] = f;
    }
}

// Spheres: block = tile of 8 (x) x 8 (y: one per warp) x 32 (z: lanes) voxels; a thread keeps its 8 x-neighbours in registers.
template <bool CALC>
__global__ void __launch_bounds__(256) sphere_fill_kernel(const float *__restrict__ g, const SphDev *__restrict__ sph, uint32_t n_sph, uint32_t res,
                                                          uint32_t tiles_y, uint32_t tiles_z, uint8_t *__restrict__ mask, float *__restrict__ field,
                                                          unsigned long long *__restrict__ ones)
{
    __shared__ SphDev s_sph[256];
    __shared__ uint32_t s_warp_hits[8];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tz = blockIdx.x % tiles_z, ty = (blockIdx.x / tiles_z) % tiles_y, tx = blockIdx.x / (tiles_z * tiles_y);
    const int32_t x0 = int32_t(tx * 8), yt = int32_t(ty * 8), zt = int32_t(tz * 32);
    const int32_t y = yt + int32_t(warp), z = zt + int32_t(lane);
    const bool live = y < int32_t(res) && z < int32_t(res);
    const float gy = g[min(y, int32_t(res) - 1)], gz = g[min(z, int32_t(res) - 1)];
    float gx[8], f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        gx[i] = g[min(x0 + i, int32_t(res) - 1)];
        f[i] = 0.f;
    }
    uint32_t in_shape = 0;

    for (uint32_t base = 0; base < n_sph; base += 256) {
        // shapes of this batch whose box meets the tile, kept in acceptance order
        const uint32_t i = base + threadIdx.x;
        SphDev mine;
        bool hit = false;
        if (i < n_sph) {
            mine = sph[i];
            hit = mine.lo[0] < x0 + 8 && mine.hi[0] > x0 && mine.lo[1] < yt + 8 && mine.hi[1] > yt && mine.lo[2] < zt + 32 && mine.hi[2] > zt;
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, hit);
        __syncthreads(); // previous batch fully consumed
        if (lane == 0) s_warp_hits[warp] = __popc(vote);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < 8; w++) {
            const uint32_t c = s_warp_hits[w];
            before += w < warp ? c : 0;
            total += c;
        }
        if (hit) s_sph[before + __popc(vote & ((1u << lane) - 1u))] = mine;
        __syncthreads();

        for (uint32_t j = 0; j < total; j++) {
            const SphDev &s = s_sph[j];
            if (!live || y < s.lo[1] || y >= s.hi[1] || z < s.lo[2] || z >= s.hi[2]) continue;
            const float p1 = __fsub_rn(gy, s.cy), p2 = __fsub_rn(gz, s.cz);
            const float s1 = sq_rn(p1), s2 = sq_rn(p2);
            const int32_t lo0 = s.lo[0] - x0, hi0 = s.hi[0] - x0;
            const float cx = s.cx, r2 = s.r2;
            const double amp = s.amp;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < lo0 || i >= hi0) continue;
                const float p0 = __fsub_rn(gx[i], cx);
                const float d2 = __fadd_rn(__fadd_rn(sq_rn(p0), s1), s2);
                if (d2 <= r2) in_shape |= 1u << i;
                if (CALC) {
                    double term = 0.0; // inside: the reference adds 0.f
                    if (d2 > r2) {
                        const float cos2 = __fdiv_rn(s2, d2); // (p . B0)^2 / d2 with B0 = (0,0,1)
                        term = __dmul_rn(__ddiv_rn(__ddiv_rn(amp, (double)d2), (double)__fsqrt_rn(d2)), __dsub_rn((double)cos2, 1. / 3.));
                    }
                    f[i] = __double2float_rn(__dadd_rn((double)f[i], term));
                }
            }
        }
    }

    uint32_t n1 = 0;
    if (live) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (x0 + i >= int32_t(res)) break;
            const uint64_t p = (uint64_t(x0 + i) * res + uint32_t(y)) * res + uint32_t(z);
            const uint32_t bit = (in_shape >> i) & 1u;
            mask[p] = uint8_t(bit);
            if (CALC) field[p] = f[i];
            n1 += bit;
        }
    }
    n1 = __reduce_add_sync(0xffffffffu, n1);
    if (lane == 0 && n1) atomicAdd(ones, (unsigned long long)n1);
}

// ---------------------------------------------------------------- host driver ----------------------------------------------------------------

struct FillResult {
    uint32_t n_launches = 0;
    uint64_t ones = 0;
    uint64_t exact_columns = 0;
    float kernel_ms = 0.f;
    std::string error;
};

#define SWK_PH_CK(call)                                                                    \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            out.error = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            goto done;                                                                     \
        }                                                                                  \
    } while (0)

inline int32_t box_lo(int32_t centre, int32_t reach) { return std::max(0, centre - reach); }

// Fills d_mask (and d_field when calc) on `stream`; all scratch is allocated and freed here.  Returns SWK_OK or SWK_ERR_CUDA.
inline int fill_device(const swk_phantom_spec &sp, const std::vector<Shape> &shapes, uint8_t *d_mask, float *d_field, cudaStream_t stream, int sm_count,
                       FillResult &out)
{
    const uint32_t res = uint32_t(sp.resolution);
    const uint64_t V = uint64_t(res) * res * res;
    const bool calc = d_field != nullptr;
    const float fov = sp.fov_um, h = fov / sp.resolution;
    const int32_t ires = int32_t(res);
    float *d_g = nullptr, *d_field2 = nullptr;
    void *d_shapes = nullptr;
    uint8_t *d_mask2 = nullptr;
    uint32_t *d_exact = nullptr;
    unsigned long long *d_cnt = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long cnt_h[2] = {0, 0};
    int rc = SWK_ERR_CUDA;

    SWK_PH_CK(cudaEventCreate(&ev0));
    SWK_PH_CK(cudaEventCreate(&ev1));
    SWK_PH_CK(cudaMalloc(&d_cnt, 2 * sizeof(unsigned long long)));
    SWK_PH_CK(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), stream));

    if (sp.shape == SWK_SHAPE_TWOPOOLS) { // phantom_twopools.cpp:55: the first half of the flat array
        SWK_PH_CK(cudaEventRecord(ev0, stream));
        SWK_PH_CK(cudaMemsetAsync(d_mask, 1, V / 2, stream));
        SWK_PH_CK(cudaMemsetAsync(d_mask + V / 2, 0, V - V / 2, stream));
        SWK_PH_CK(cudaEventRecord(ev1, stream));
        SWK_PH_CK(cudaStreamSynchronize(stream));
        SWK_PH_CK(cudaEventElapsedTime(&out.kernel_ms, ev0, ev1));
        out.ones = V / 2;
        rc = SWK_OK;
        goto done;
    }
    {
        const std::vector<float> g = voxel_centres(fov, res);
        SWK_PH_CK(cudaMalloc(&d_g, res * sizeof(float)));
        SWK_PH_CK(cudaMemcpyAsync(d_g, g.data(), res * sizeof(float), cudaMemcpyHostToDevice, stream));
        const uint32_t n = uint32_t(shapes.size());

        if (sp.shape == SWK_SHAPE_CYLINDER) {
            const bool force_exact = getenv("SWK_PHANTOM_FORCE_EXACT") != nullptr; // test hook for cyl_exact_kernel
            std::vector<CylDev> cyl(n);
            for (uint32_t i = 0; i < n; i++) {
                const Shape &s = shapes[i];
                CylDev &c = cyl[i];
                c.cx = s.x; c.cy = s.y; c.cz = s.z; c.r2 = s.r * s.r;
                const int32_t rv = int32_t(std::ceil(s.r / h) + 1), vx = int32_t(s.x / h), vy = int32_t(s.y / h);
                if (calc) { // field map: 20 radii around the axis (phantom_cylinder.cpp:224-230)
                    c.x0 = box_lo(vx, rv * 20); c.x1 = std::min(ires, vx + rv * 20);
                    c.y0 = box_lo(vy, rv * 20); c.y1 = std::min(ires, vy + rv * 20);
                } else {
                    c.x0 = box_lo(vx, rv); c.x1 = std::min(ires, vx + rv + 2);
                    c.y0 = box_lo(vy, rv); c.y1 = std::min(ires, vy + rv + 2);
                }
                float worst = 0.f;
                for (uint32_t z = 0; z < res; z++) {
                    const float back = (g[z] - s.z) + s.z, zr = g[z] - back;
                    worst = std::max(worst, zr * zr);
                }
                c.zres2_max = worst;
                if (force_exact) c.zres2_max = INFINITY; // fl(d2 + inf) != d2: every column outside a cylinder takes the voxel-by-voxel kernel
            }
            CylConst k{};
            { // B0 = Ry(orientation) (0,0,1) in float from double sin/cos (phantom_base.h:133-142), projected and normalised
                const float rad = sp.orientation_deg * float(0.0174532925199433);
                const float sn = float(std::sin(double(rad))), cs = float(std::cos(double(rad)));
                float b[3] = {cs * 0.f + sn * 1.f, 0.f, 0.f};
                const float len = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
                if (len != 0) { b[0] /= len; b[1] /= len; }
                k.b0x = b[0]; k.b0y = b[1];
                const float ct = float(std::cos(sp.orientation_deg * M_PI / 180));
                const float ct2 = ct * ct;
                k.sin2 = float(1. - ct2);
                k.k_out = 2 * M_PI * (1 - sp.oxy_level) * sp.dchi;
                k.k_in = 2 * M_PI * (1 - sp.oxy_level) * sp.dchi * (ct2 - 1.0 / 3.0);
            }
            const uint64_t cols = uint64_t(res) * res;
            SWK_PH_CK(cudaMalloc(&d_shapes, std::max<size_t>(1, n) * sizeof(CylDev)));
            SWK_PH_CK(cudaMemcpyAsync(d_shapes, cyl.data(), n * sizeof(CylDev), cudaMemcpyHostToDevice, stream));
            SWK_PH_CK(cudaMalloc(&d_mask2, cols));
            SWK_PH_CK(cudaMalloc(&d_field2, cols * sizeof(float)));
            SWK_PH_CK(cudaMalloc(&d_exact, cols * sizeof(uint32_t)));
            unsigned int *cnt32 = reinterpret_cast<unsigned int *>(d_cnt);
            const uint32_t slab_blocks = uint32_t((cols + 255) / 256), bc_blocks = uint32_t(std::min<uint64_t>(uint64_t(sm_count) * 8, V / 4096 + 1));
            SWK_PH_CK(cudaEventRecord(ev0, stream));
            // TMA bulk stores from shared memory by default (measured 0.855 ms vs 0.975 ms for the 1000^3 phantom); SWK_PHANTOM_BCAST=stg selects
            // the plain 16 B store kernel for A/B runs
            const char *bc = getenv("SWK_PHANTOM_BCAST");
            const bool bulk = !(bc && std::string(bc) == "stg");
            const uint32_t bulk_blocks = uint32_t(std::min<uint64_t>(uint64_t(sm_count) * 5, V / 4096 + 1));
            if (calc) {
                cyl_slab_kernel<true><<<slab_blocks, 256, 0, stream>>>(d_g, static_cast<const CylDev *>(d_shapes), n, res, k, d_mask2, d_field2, d_exact, cnt32);
                if (bulk) slab_broadcast_bulk_kernel<true><<<bulk_blocks, 256, 40960, stream>>>(d_mask2, d_field2, res, V, d_mask, d_field);
                else slab_broadcast_kernel<true><<<bc_blocks, 256, 0, stream>>>(d_mask2, d_field2, res, V, d_mask, d_field);
            } else {
                cyl_slab_kernel<false><<<slab_blocks, 256, 0, stream>>>(d_g, static_cast<const CylDev *>(d_shapes), n, res, k, d_mask2, d_field2, d_exact, cnt32);
                if (bulk) slab_broadcast_bulk_kernel<false><<<bulk_blocks, 256, 40960, stream>>>(d_mask2, d_field2, res, V, d_mask, d_field);
                else slab_broadcast_kernel<false><<<bc_blocks, 256, 0, stream>>>(d_mask2, d_field2, res, V, d_mask, d_field);
            }
            out.n_launches = 2;
            SWK_PH_CK(cudaGetLastError());
            unsigned int c32[2] = {0, 0};
            SWK_PH_CK(cudaMemcpyAsync(c32, d_cnt, sizeof c32, cudaMemcpyDeviceToHost, stream));
            SWK_PH_CK(cudaStreamSynchronize(stream));
            if (c32[1]) {
                cyl_exact_kernel<<<c32[1], 256, 0, stream>>>(d_g, static_cast<const CylDev *>(d_shapes), n, res, k, d_exact, d_field);
                SWK_PH_CK(cudaGetLastError());
                out.n_launches++;
            }
            SWK_PH_CK(cudaEventRecord(ev1, stream));
            SWK_PH_CK(cudaStreamSynchronize(stream));
            SWK_PH_CK(cudaEventElapsedTime(&out.kernel_ms, ev0, ev1));
            out.ones = uint64_t(c32[0]) * res;
            out.exact_columns = c32[1];
        } else {
            std::vector<SphDev> sph(n);
            const double k = 4 * M_PI * (1 - sp.oxy_level) * sp.dchi;
            for (uint32_t i = 0; i < n; i++) {
                const Shape &s = shapes[i];
                SphDev &d = sph[i];
                d.cx = s.x; d.cy = s.y; d.cz = s.z; d.r2 = s.r * s.r;
                d.amp = k * d.r2 * s.r;
                const int32_t rv = int32_t(std::ceil(s.r / h) + 1);
                const float c[3] = {s.x, s.y, s.z};
                for (int a = 0; a < 3; a++) {
                    const int32_t v = int32_t(c[a] / h);
                    if (calc) { d.lo[a] = box_lo(v, rv * 20); d.hi[a] = std::min(ires, v + rv * 20); }
                    else { d.lo[a] = box_lo(v, rv); d.hi[a] = std::min(ires, v + rv + 2); }
                }
            }
            SWK_PH_CK(cudaMalloc(&d_shapes, std::max<size_t>(1, n) * sizeof(SphDev)));
            SWK_PH_CK(cudaMemcpyAsync(d_shapes, sph.data(), n * sizeof(SphDev), cudaMemcpyHostToDevice, stream));
            const uint32_t tiles_x = (res + 7) / 8, tiles_y = (res + 7) / 8, tiles_z = (res + 31) / 32;
            const uint64_t blocks = uint64_t(tiles_x) * tiles_y * tiles_z;
            if (blocks > 0x7fffffffull) { out.error = "phantom too large for one launch"; rc = SWK_ERR_INVALID; goto done; }
            SWK_PH_CK(cudaEventRecord(ev0, stream));
            if (calc)
                sphere_fill_kernel<true><<<uint32_t(blocks), 256, 0, stream>>>(d_g, static_cast<const SphDev *>(d_shapes), n, res, tiles_y, tiles_z, d_mask, d_field, d_cnt);
            else
                sphere_fill_kernel<false><<<uint32_t(blocks), 256, 0, stream>>>(d_g, static_cast<const SphDev *>(d_shapes), n, res, tiles_y, tiles_z, d_mask, d_field, d_cnt);
            out.n_launches = 1;
            SWK_PH_CK(cudaGetLastError());
            SWK_PH_CK(cudaEventRecord(ev1, stream));
            SWK_PH_CK(cudaMemcpyAsync(cnt_h, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
            SWK_PH_CK(cudaStreamSynchronize(stream));
            SWK_PH_CK(cudaEventElapsedTime(&out.kernel_ms, ev0, ev1));
            out.ones = cnt_h[0];
        }
        rc = SWK_OK;
    }
done:
    if (d_g) cudaFree(d_g);
    if (d_shapes) cudaFree(d_shapes);
    if (d_mask2) cudaFree(d_mask2);
    if (d_field2) cudaFree(d_field2);
    if (d_exact) cudaFree(d_exact);
    if (d_cnt) cudaFree(d_cnt);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    return rc;
}
#undef SWK_PH_CK

// placement for any shape; place_ms = host time
inline int place(const swk_phantom_spec &sp, std::vector<Shape> &shapes, float &place_ms, std::string &err)
{
    shapes.clear();
    place_ms = 0.f;
    if (sp.shape != SWK_SHAPE_CYLINDER && sp.shape != SWK_SHAPE_SPHERE && sp.shape != SWK_SHAPE_TWOPOOLS) { err = "unknown phantom shape"; return SWK_ERR_INVALID; }
    if (!(sp.fov_um > 0.f) || sp.resolution == 0) { err = "FOV or resolution is not set"; return SWK_ERR_INVALID; } // phantom_base.cpp:110-114
    if (sp.resolution > 2048) { err = "resolution above 2048 is not supported"; return SWK_ERR_INVALID; }
    if (sp.shape == SWK_SHAPE_TWOPOOLS) return SWK_OK;
    if (!(sp.volume_fraction > 0.f)) { err = "volume fraction must be positive"; return SWK_ERR_INVALID; }
    const auto t0 = std::chrono::steady_clock::now();
    int rc;
    if (sp.shape == SWK_SHAPE_CYLINDER) rc = place_cylinders(sp, voxel_centres(sp.fov_um, sp.resolution), shapes);
    else rc = place_spheres(sp, shapes);
    place_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc == SWK_PLACE_STALLED) {
        err = "shape placement does not converge for this FoV / resolution / radius / volume fraction (the reference would loop forever)";
        return SWK_ERR_INVALID;
    }
    if (rc != SWK_OK) err = "the radius of the shapes is too large for the given FOV"; // phantom_cylinder.cpp:87-91
    return rc;
}

inline bool wants_fieldmap(const swk_phantom_spec &sp) { return sp.shape != SWK_SHAPE_TWOPOOLS && sp.oxy_level >= 0; }

} // namespace phantom
} // namespace swk
