// spinwalk_b200/csrc/phantom_mesh.cuh — triangle-mesh phantom (`spinwalk phantom -p`), voxelised on the GPU.
//
// Reference: src/phantom/phantom_ply.cpp:141-227.  A voxel centre is inside the mesh when the ray (1,0,0) from it hits an odd
// number of triangles (Möller-Trumbore in mixed double / float arithmetic, :90-111).  The reference walks a median-split BVH
// per voxel on the host (one z slice after another); its boxes prune on (y, z) only (phantom_ply.h:37-42), so a triangle is
// tested for a ray exactly when the ray's (y, z) lies inside the box of the LEAF that holds it (a leaf's box lies inside all
// its ancestors').  Here:
//   host    the same mesh transform (mm -> um, centred in the FoV) and the same BVH build — std::sort with the reference's
//           centroid comparator, so even ties fall the same way — only to give every triangle its leaf box; then the
//           triangles are binned by the (y, z) voxel rows their leaf box covers (CSR), because a +x ray never leaves its row;
//   device  mesh_fill_kernel: one thread per voxel (z fastest => coalesced mask bytes) runs the reference's hit test over its
//           row's candidates, every product and sum issued in the reference's type and order with round-to-nearest
//           intrinsics (no FMA contraction) => the mask is bit-identical.
// Bound: FP64 issue (about 30 double operations per voxel-candidate pair); the mask itself is 1 B per voxel.
#pragma once

#include <algorithm>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "phantom.cuh"

namespace swk {
namespace phantom {

struct MeshTri {
    double v0[3], e1[3], e2[3], h[3]; // h = cross(dir, e2) with dir = (1,0,0), evaluated like the reference (so h[0] is a signed zero)
    double ymin, ymax, zmin, zmax;    // box of the BVH leaf holding the triangle
    float f;                          // 1.0f / a
    int32_t valid;                    // |a| >= 1e-6f (not parallel to the ray)
};

struct HostTri { double v0[3], v1[3], v2[3]; double ymin, ymax, zmin, zmax; };

inline void tri_bounds(const std::vector<HostTri> &t, size_t start, size_t end, double mn[3], double mx[3])
{ // computeAABB (phantom_ply.cpp:43-52)
    for (int k = 0; k < 3; k++) { mn[k] = std::numeric_limits<double>::max(); mx[k] = std::numeric_limits<double>::lowest(); }
    for (size_t i = start; i < end; i++)
        for (const double *v : {t[i].v0, t[i].v1, t[i].v2})
            for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], v[k]); mx[k] = std::max(mx[k], v[k]); }
}

// buildBVH (phantom_ply.cpp:73-86) reduced to what the traversal needs from it: every triangle's leaf box
inline void assign_leaf_boxes(std::vector<HostTri> &t, size_t start, size_t end)
{
    double mn[3], mx[3];
    tri_bounds(t, start, end, mn, mx);
    if (end - start <= 4) {
        for (size_t i = start; i < end; i++) { t[i].ymin = mn[1]; t[i].ymax = mx[1]; t[i].zmin = mn[2]; t[i].zmax = mx[2]; }
        return;
    }
    const double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    int axis = 0; // longest axis, with the reference's tie rules (:57-60)
    if (ey > ex) axis = 1;
    if (ez > std::max(ex, ey)) axis = 2;
    std::sort(t.begin() + start, t.begin() + end, [axis](const HostTri &a, const HostTri &b) {
        const double ca = (a.v0[axis] + a.v1[axis] + a.v2[axis]) / 3.0f, cb = (b.v0[axis] + b.v1[axis] + b.v2[axis]) / 3.0f;
        return ca < cb;
    });
    const size_t mid = start + (end - start) / 2;
    assign_leaf_boxes(t, start, mid);
    assign_leaf_boxes(t, mid, end);
}

// rayIntersectsTriangle (phantom_ply.cpp:90-111) for the ray (1,0,0) from o
__device__ __forceinline__ bool mesh_ray_hits(const double o[3], const MeshTri &t)
{
    const double sx = __dsub_rn(o[0], t.v0[0]), sy = __dsub_rn(o[1], t.v0[1]), sz = __dsub_rn(o[2], t.v0[2]);
    const float u = __fmul_rn(t.f, __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(sx, t.h[0]), __dmul_rn(sy, t.h[1])), __dmul_rn(sz, t.h[2]))));
    if (u < 0.0f || u > 1.0f) return false;
    const double qx = __dsub_rn(__dmul_rn(sy, t.e1[2]), __dmul_rn(sz, t.e1[1]));
    const double qy = __dsub_rn(__dmul_rn(sz, t.e1[0]), __dmul_rn(sx, t.e1[2]));
    const double qz = __dsub_rn(__dmul_rn(sx, t.e1[1]), __dmul_rn(sy, t.e1[0]));
    const float v = __fmul_rn(t.f, __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(1., qx), __dmul_rn(0., qy)), __dmul_rn(0., qz))));
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    const float tt = __fmul_rn(t.f, __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(t.e2[0], qx), __dmul_rn(t.e2[1], qy)), __dmul_rn(t.e2[2], qz))));
    return tt >= 0.0f;
}

// one thread per voxel, z fastest.  row = py * res + pz indexes the CSR of candidate triangles.
__global__ void __launch_bounds__(256) mesh_fill_kernel(const float *__restrict__ g, const MeshTri *__restrict__ tri, const uint32_t *__restrict__ row_start,
                                                        const uint32_t *__restrict__ row_items, uint32_t res, uint64_t V, double bx0, double bx1, double by0, double by1,
                                                        double bz0, double bz1, uint8_t *__restrict__ mask, unsigned long long *__restrict__ ones)
{
    const uint64_t p = uint64_t(blockIdx.x) * 256 + threadIdx.x;
    uint32_t inside = 0;
    if (p < V) {
        const uint32_t pz = uint32_t(p % res), py = uint32_t((p / res) % res), px = uint32_t(p / (uint64_t(res) * res));
        const double o[3] = {double(g[px]), double(g[py]), double(g[pz])};
        // outside the mesh's bounding box: stays 0 (phantom_ply.cpp:206-207)
        if (!(o[0] < bx0 || o[0] > bx1 || o[1] < by0 || o[1] > by1 || o[2] < bz0 || o[2] > bz1)) {
            const uint32_t row = py * res + pz;
            uint32_t hits = 0;
            for (uint32_t k = row_start[row]; k < row_start[row + 1]; k++) {
                const MeshTri &t = tri[row_items[k]];
                if (o[1] < t.ymin || o[1] > t.ymax || o[2] < t.zmin || o[2] > t.zmax) continue; // the leaf box test (AABB::intersectsRay)
                if (!t.valid) continue;
                hits += mesh_ray_hits(o, t) ? 1u : 0u;
            }
            inside = hits & 1u;
        }
        mask[p] = uint8_t(inside);
    }
    const unsigned int n1 = __syncthreads_count(inside != 0);
    if (threadIdx.x == 0 && n1) atomicAdd(ones, (unsigned long long)n1);
}

struct MeshResult {
    uint64_t ones = 0, row_items = 0;
    float kernel_ms = 0.f, prep_ms = 0.f;
    std::string error;
};

// vertices: double [nv][3] in the PLY file's unit (mm); faces: [nf][3].  Fills d_mask [res]^3 on `stream`.
inline int fill_mesh_device(float fov_um, uint32_t res, const double *vertices, uint64_t n_vertices, const uint64_t *faces, uint64_t n_faces, uint8_t *d_mask,
                            cudaStream_t stream, MeshResult &out)
{
    const auto t_host = std::chrono::steady_clock::now();
    const uint64_t V = uint64_t(res) * res * res;
    // ---- the reference's mesh transform (phantom_ply.cpp:158-180)
    std::vector<HostTri> tris(n_faces);
    for (uint64_t i = 0; i < n_faces; i++) {
        const uint64_t *f = faces + 3 * i;
        if (f[0] >= n_vertices || f[1] >= n_vertices || f[2] >= n_vertices) { out.error = "face index out of range"; return SWK_ERR_INVALID; }
        for (int k = 0; k < 3; k++) {
            tris[i].v0[k] = vertices[3 * f[0] + k] * 1e3;
            tris[i].v1[k] = vertices[3 * f[1] + k] * 1e3;
            tris[i].v2[k] = vertices[3 * f[2] + k] * 1e3;
        }
    }
    double mn[3], mx[3];
    tri_bounds(tris, 0, tris.size(), mn, mx);
    const double half = fov_um / 2.;
    double shift[3];
    for (int k = 0; k < 3; k++) shift[k] = (mx[k] + mn[k]) / 2.;
    for (HostTri &t : tris)
        for (double *v : {t.v0, t.v1, t.v2})
            for (int k = 0; k < 3; k++) v[k] = v[k] + half - shift[k];
    assign_leaf_boxes(tris, 0, tris.size());
    tri_bounds(tris, 0, tris.size(), mn, mx); // root->bounds

    // ---- per-triangle constants + rows covered by the leaf box
    const std::vector<float> g = voxel_centres(fov_um, res);
    std::vector<double> gd(g.begin(), g.end());
    std::vector<MeshTri> dev(n_faces);
    std::vector<uint32_t> y0(n_faces), y1(n_faces), z0(n_faces), z1(n_faces);
    std::vector<uint32_t> row_start(size_t(res) * res + 1, 0);
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_faces; i++) {
        const HostTri &s = tris[i];
        MeshTri &d = dev[i];
        const double dir[3] = {1., 0., 0.};
        for (int k = 0; k < 3; k++) { d.v0[k] = s.v0[k]; d.e1[k] = s.v1[k] - s.v0[k]; d.e2[k] = s.v2[k] - s.v0[k]; }
        d.h[0] = dir[1] * d.e2[2] - dir[2] * d.e2[1];
        d.h[1] = dir[2] * d.e2[0] - dir[0] * d.e2[2];
        d.h[2] = dir[0] * d.e2[1] - dir[1] * d.e2[0];
        const float a = float(d.e1[0] * d.h[0] + d.e1[1] * d.h[1] + d.e1[2] * d.h[2]);
        d.valid = !(std::abs(a) < 1e-6f);
        d.f = 1.0f / a;
        d.ymin = s.ymin; d.ymax = s.ymax; d.zmin = s.zmin; d.zmax = s.zmax;
        // voxel rows whose centre passes `!(o < min || o > max)`: centres ascend, so two binary searches per axis
        y0[i] = uint32_t(std::lower_bound(gd.begin(), gd.end(), s.ymin) - gd.begin());
        y1[i] = uint32_t(std::upper_bound(gd.begin(), gd.end(), s.ymax) - gd.begin());
        z0[i] = uint32_t(std::lower_bound(gd.begin(), gd.end(), s.zmin) - gd.begin());
        z1[i] = uint32_t(std::upper_bound(gd.begin(), gd.end(), s.zmax) - gd.begin());
        if (!d.valid || y1[i] <= y0[i] || z1[i] <= z0[i]) { y1[i] = y0[i]; continue; }
        for (uint32_t y = y0[i]; y < y1[i]; y++)
            for (uint32_t z = z0[i]; z < z1[i]; z++) row_start[size_t(y) * res + z + 1]++;
        total += uint64_t(y1[i] - y0[i]) * (z1[i] - z0[i]);
    }
    if (total > 1500ull * 1000 * 1000) {
        out.error = "mesh too large for the row binning (more than 1.5e9 triangle-row pairs): raise the mesh resolution or lower the phantom resolution";
        return SWK_ERR_MEMORY;
    }
    for (size_t r = 0; r < size_t(res) * res; r++) row_start[r + 1] += row_start[r];
    std::vector<uint32_t> items(std::max<uint64_t>(1, total));
    {
        std::vector<uint32_t> cursor(row_start.begin(), row_start.end() - 1);
        for (uint64_t i = 0; i < n_faces; i++)
            for (uint32_t y = y0[i]; y < y1[i]; y++)
                for (uint32_t z = z0[i]; z < z1[i]; z++) items[cursor[size_t(y) * res + z]++] = uint32_t(i);
    }
    out.row_items = total;
    out.prep_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_host).count();

    float *d_g = nullptr;
    MeshTri *d_tri = nullptr;
    uint32_t *d_start = nullptr, *d_items = nullptr;
    unsigned long long *d_ones = nullptr, ones_h = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rc = SWK_ERR_CUDA;
#define SWK_MESH_CK(call)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            out.error = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            goto done;                                                                     \
        }                                                                                  \
    } while (0)
    SWK_MESH_CK(cudaEventCreate(&ev0));
    SWK_MESH_CK(cudaEventCreate(&ev1));
    SWK_MESH_CK(cudaMalloc(&d_g, res * sizeof(float)));
    SWK_MESH_CK(cudaMalloc(&d_tri, std::max<size_t>(1, n_faces) * sizeof(MeshTri)));
    SWK_MESH_CK(cudaMalloc(&d_start, row_start.size() * sizeof(uint32_t)));
    SWK_MESH_CK(cudaMalloc(&d_items, items.size() * sizeof(uint32_t)));
    SWK_MESH_CK(cudaMalloc(&d_ones, sizeof(unsigned long long)));
    SWK_MESH_CK(cudaMemcpyAsync(d_g, g.data(), res * sizeof(float), cudaMemcpyHostToDevice, stream));
    SWK_MESH_CK(cudaMemcpyAsync(d_tri, dev.data(), n_faces * sizeof(MeshTri), cudaMemcpyHostToDevice, stream));
    SWK_MESH_CK(cudaMemcpyAsync(d_start, row_start.data(), row_start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    SWK_MESH_CK(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    SWK_MESH_CK(cudaMemsetAsync(d_ones, 0, sizeof(unsigned long long), stream));
    SWK_MESH_CK(cudaEventRecord(ev0, stream));
    mesh_fill_kernel<<<uint32_t((V + 255) / 256), 256, 0, stream>>>(d_g, d_tri, d_start, d_items, res, V, mn[0], mx[0], mn[1], mx[1], mn[2], mx[2], d_mask, d_ones);
    SWK_MESH_CK(cudaGetLastError());
    SWK_MESH_CK(cudaEventRecord(ev1, stream));
    SWK_MESH_CK(cudaMemcpyAsync(&ones_h, d_ones, sizeof ones_h, cudaMemcpyDeviceToHost, stream));
    SWK_MESH_CK(cudaStreamSynchronize(stream));
    SWK_MESH_CK(cudaEventElapsedTime(&out.kernel_ms, ev0, ev1));
    out.ones = ones_h;
    rc = SWK_OK;
#undef SWK_MESH_CK
done:
    if (d_g) cudaFree(d_g);
    if (d_tri) cudaFree(d_tri);
    if (d_start) cudaFree(d_start);
    if (d_items) cudaFree(d_items);
    if (d_ones) cudaFree(d_ones);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    return rc;
}

} // namespace phantom
} // namespace swk
