/* include/spinwalk_phantom.h — C-ABI of the B200 phantom generator (part of libspinwalk_b200.so).
 *
 * the producer on the input side of the `sim` hot path.  Replaces
 *      src/phantom/handler.cpp:10-35                  phantom::handler::execute (what `spinwalk phantom` calls)
 *      src/phantom/phantom_base.cpp:107-143           voxel-centre grid (12 B per voxel on the host in the reference)
 *      src/phantom/phantom_cylinder.cpp:85-130        random placement of parallel cylinders (sequential host RNG)
 *      src/phantom/phantom_cylinder.cpp:184-275       mask + analytic dB of infinite cylinders   -> CUDA kernels
 *      src/phantom/phantom_sphere.cpp:79-119          random placement of spheres
 *      src/phantom/phantom_sphere.cpp:121-198         mask + dipole field of spheres             -> CUDA kernel
 *      src/phantom/phantom_twopools.cpp:40-63         half/half mask
 *      src/phantom/phantom_ply.cpp:141-227            closed triangle mesh -> mask (ray parity)  -> CUDA kernel (swk_phantom_mesh)
 * The shape placement is inherently sequential (every accepted shape changes the acceptance test of the next) and stays on
 * the host with the reference's engines (std::mt19937 / std::minstd_rand, uniform_real_distribution<float>); the O(V x shapes)
 * voxel fill runs on the device, with the reference's float/double expression order, so that the mask and the field map
 * are bit-identical to a serial x86-64 build of the reference (tests/test_phantom_gpu.py).
 *
 * Conventions as in spinwalk_engine.h: plain C, int status (SWK_OK == 0), caller owns host buffers, no CPU fallback for
 * the voxel fill (swk_phantom_generate fails without a CUDA device; swk_phantom_shapes is host-only by nature).
 */
#ifndef SPINWALK_PHANTOM_H
#define SPINWALK_PHANTOM_H

#include <stddef.h>
#include <stdint.h>

#include "spinwalk_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

enum swk_phantom_shape { SWK_SHAPE_CYLINDER = 0, SWK_SHAPE_SPHERE = 1, SWK_SHAPE_TWOPOOLS = 2 };

/* ≙ phantom::execute_args (src/phantom/handler.h:10-25) = the options of `spinwalk phantom` (src/spinwalk.cpp:58-72) */
typedef struct swk_phantom_spec {
    int32_t  shape;            /* -c / -s / -t                                                                     */
    float    fov_um;           /* -f  isotropic field of view, µm                                                  */
    uint64_t resolution;       /* -z  voxels per axis                                                              */
    float    dchi;             /* -d  susceptibility difference (default 0.11e-6)                                  */
    float    oxy_level;        /* -y  Y; < 0 => mask only, no field map (phantom_base.cpp:56)                      */
    float    radius_um;        /* -r  < 0 => random radius below |r|                                               */
    float    volume_fraction;  /* -v  target volume fraction, percent                                              */
    float    orientation_deg;  /* -n  cylinder axis vs B0 (cylinders only)                                         */
    int32_t  seed;             /* -e  < 0 => std::random_device (phantom_base.cpp:26,40-41)                        */
} swk_phantom_spec;

typedef struct swk_phantom_stats {
    uint32_t n_shapes;         /* cylinders / spheres placed                                                       */
    float    volume_fraction;  /* actual volume fraction in percent = the `bvf` dataset (phantom_cylinder.cpp:270)  */
    float    place_ms;         /* host time of the placement loop                                                  */
    float    kernel_ms;        /* device time of the voxel fill (CUDA events on the launching stream)              */
    uint32_t n_launches;       /* kernels launched                                                                 */
    uint64_t exact_columns;    /* cylinders: (x,y) columns re-evaluated voxel by voxel because the reference's      */
                               /* z-residual could change a rounding (normally 0; see csrc/phantom.cuh)             */
} swk_phantom_stats;

/* Placement only (host): shapes [cap][4] = centre x, y, z and radius in µm, in acceptance order.  *n_shapes is the number
 * placed even when it exceeds cap.  SWK_ERR_INVALID when the reference refuses (2*radius >= fov, phantom_cylinder.cpp:87). */
int swk_phantom_shapes(const swk_phantom_spec *spec, float *shapes, uint32_t cap, uint32_t *n_shapes);

/* Placement + voxel fill on `device`.  mask: uint8 [res][res][res], x slowest (the /mask dataset); fieldmap_T: float, same
 * shape, Tesla at B0 = 1 T (the /fieldmap dataset), required iff oxy_level >= 0 and shape != two pools.  The pointers are
 * host pointers, or device pointers on `device` when on_device != 0.  stats may be NULL. */
int swk_phantom_generate(int device, const swk_phantom_spec *spec, uint8_t *mask, float *fieldmap_T, int on_device,
                         swk_phantom_stats *stats);

/* The same, straight into an engine: the phantom is generated in the engine's own device buffers and becomes its current
 * phantom (≙ swk_set_phantom with fov = fov_um * 1e-6 on every axis, phantom_base.cpp:63-66) without touching the host. */
int swk_generate_phantom(swk_engine *e, const swk_phantom_spec *spec, swk_phantom_stats *stats);

/* Copies the engine's current phantom to host buffers (either may be NULL), e.g. to save a generated phantom. */
int swk_get_phantom(swk_engine *e, uint8_t *mask, float *fieldmap_T);

/* Triangle-mesh phantom (≙ phantom::ply::run(false), src/phantom/phantom_ply.cpp:141-227, `spinwalk phantom -p -i mesh.ply`).
 * vertices: double [n_vertices][3] in the PLY file's unit (mm; converted to µm and centred in the FoV like the reference, :160-175);
 * faces: [n_faces][3] vertex indices (triangles only).  mask: uint8 [res][res][res], x slowest, 1 where the +x ray from the voxel
 * centre hits an odd number of triangles — bit-identical to the reference (same BVH leaf boxes, same mixed double/float hit test).
 * No field map (the reference forces Y = -1).  stats->n_shapes = n_faces; stats->volume_fraction is the actual percentage (the
 * reference leaves the `bvf` dataset of a mesh phantom at 0; the CLI writes 0 as well). */
int swk_phantom_mesh(int device, float fov_um, uint64_t resolution, const double *vertices, uint64_t n_vertices, const uint64_t *faces,
                     uint64_t n_faces, uint8_t *mask, int on_device, swk_phantom_stats *stats);

const char *swk_phantom_last_error(void); /* message of the last failed swk_phantom_* call on this thread */

#ifdef __cplusplus
}
#endif
#endif /* SPINWALK_PHANTOM_H */
