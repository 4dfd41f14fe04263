/* oracle/phantom_oracle.h — TEST INFRASTRUCTURE: C restatement of the reference's phantom generators (see phantom_oracle.c). */
#ifndef SWO_PHANTOM_ORACLE_H
#define SWO_PHANTOM_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { SWO_SHAPE_CYLINDER = 0, SWO_SHAPE_SPHERE = 1, SWO_SHAPE_TWOPOOLS = 2 };

/* the arguments of `spinwalk phantom` (src/spinwalk.cpp:58-72, src/phantom/handler.h:10-25) */
typedef struct swo_phantom_spec {
    int32_t  shape;
    float    fov_um;          /* -f */
    uint64_t resolution;      /* -z */
    float    dchi;            /* -d */
    float    Y;               /* -y; < 0 => mask only */
    float    radius_um;       /* -r; < 0 => random radius below |r| */
    float    volume_fraction; /* -v, percent */
    float    orientation_deg; /* -n (cylinders) */
    int32_t  seed;            /* -e, >= 0 */
} swo_phantom_spec;

/* shapes: [cap][4] = centre x, y, z and radius in µm.  Returns 0, or 1 when the reference would refuse (radius too large). */
int swo_phantom_shapes(const swo_phantom_spec *s, float *shapes, uint32_t cap, uint32_t *n_shapes);
/* mask: uint8 [res][res][res] x slowest; fieldmap: float, same shape (Tesla at 1 T), required when Y >= 0 */
int swo_phantom_generate(const swo_phantom_spec *s, uint8_t *mask, float *fieldmap, float *bvf, float *shapes, uint32_t cap, uint32_t *n_shapes);

/* the same, restricted to the z slices [zlo, zhi): outputs are [res][res][zhi-zlo] (full-size spot checks in seconds) */
int swo_phantom_generate_window(const swo_phantom_spec *s, int32_t zlo, int32_t zhi, uint8_t *mask, float *fieldmap, float *shapes, uint32_t cap,
                                uint32_t *n_shapes);

/* `spinwalk phantom -p`: mask of a closed triangle mesh centred in the FoV.  vertices double [nv][3] in the PLY file's unit (mm),
 * faces uint64 [nf][3].  Returns 0, 1 (no FoV / resolution) or 2 (face index out of range). */
int swo_phantom_mesh(float fov_um, uint64_t resolution, const double *vertices, uint64_t n_vertices, const uint64_t *faces, uint64_t n_faces, uint8_t *mask);

#ifdef __cplusplus
}
#endif
#endif
