// oracle/ref_gen_harness.cpp — TEST INFRASTRUCTURE.  C entry points into the reference's own OFFLINE GENERATORS, compiled
// unmodified where they lie (oracle/Makefile -> oracle/_ref/libswref_gen.so):
//   src/phantom/phantom_{base,cylinder,sphere,twopools}.cpp   `spinwalk phantom -c|-s|-t`
//   src/dwi/{pgse,handler}.cpp                                 `spinwalk dwi`
//   src/config/{config_generator,handler}.cpp                  `spinwalk config`
// Used by tests/ to pin oracle/phantom_oracle.c and the host-side generators of host/, and by bench.py's phantom CPU baseline.
// `private` is opened for this translation unit only, to read the generated shape lists out of the reference objects.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <boost/log/trivial.hpp>

#define private public
#define protected public
#include "phantom/phantom_cylinder.h"
#include "phantom/phantom_ply.h"
#include "phantom/phantom_sphere.h"
#include "phantom/phantom_twopools.h"
#undef private
#undef protected
#include "config/handler.h"
#include "dwi/handler.h"
#include "ini.h" // the reference's vendored mINI (include/ini.h), built with MINI_CASE_SENSITIVE like the reference (CMakeLists.txt:36)

namespace {
// the generators draw progress bars and banners on std::cout; keep the caller's stdout clean
struct quiet_cout {
    std::ostringstream sink;
    std::streambuf *old;
    quiet_cout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~quiet_cout() { std::cout.rdbuf(old); }
};
template <class P>
int export_phantom(P &ph, size_t res, bool has_field, uint8_t *mask, float *fieldmap, float *bvf)
{
    const size_t V = res * res * res;
    if (ph.m_mask.size() != V) return 2;
    if (mask) memcpy(mask, ph.m_mask.data(), V);
    if (fieldmap && has_field) {
        if (ph.m_fieldmap.size() != V) return 3;
        memcpy(fieldmap, ph.m_fieldmap.data(), V * sizeof(float));
    }
    if (bvf) *bvf = ph.m_volume_fraction;
    return 0;
}
} // namespace

extern "C" {

// shape: 0 cylinder, 1 sphere, 2 two pools.  shapes: [cap][4] = x, y, z, radius (µm) of every placed shape.
// Returns 0 on success, 1 if the reference's run() returned false.
int swref_phantom(int shape, float fov_um, uint64_t resolution, float dchi, float Y, float radius_um, float volume_fraction,
                  float orientation_deg, int32_t seed, uint8_t *mask, float *fieldmap, float *bvf, float *shapes, uint32_t cap,
                  uint32_t *n_shapes)
{
    quiet_cout q;
    if (n_shapes) *n_shapes = 0;
    if (shape == 0) {
        phantom::cylinder c(fov_um, resolution, dchi, Y, radius_um, volume_fraction, orientation_deg, seed, "unused.h5");
        if (!c.run(false)) return 1;
        const uint32_t n = (uint32_t)c.m_cylinder_radii.size();
        if (n_shapes) *n_shapes = n;
        for (uint32_t i = 0; shapes && i < n && i < cap; i++) {
            for (int k = 0; k < 3; k++) shapes[4 * i + k] = c.m_cylinder_points[i][k];
            shapes[4 * i + 3] = c.m_cylinder_radii[i];
        }
        return export_phantom(c, resolution, Y >= 0, mask, fieldmap, bvf);
    }
    if (shape == 1) {
        phantom::sphere s(fov_um, resolution, dchi, Y, radius_um, volume_fraction, seed, "unused.h5");
        if (!s.run(false)) return 1;
        const uint32_t n = (uint32_t)s.m_sphere_radii.size();
        if (n_shapes) *n_shapes = n;
        for (uint32_t i = 0; shapes && i < n && i < cap; i++) {
            for (int k = 0; k < 3; k++) shapes[4 * i + k] = s.m_sphere_points[i][k];
            shapes[4 * i + 3] = s.m_sphere_radii[i];
        }
        return export_phantom(s, resolution, Y >= 0, mask, fieldmap, bvf);
    }
    if (shape == 2) {
        phantom::twopools t(fov_um, resolution, "unused.h5");
        if (!t.run(false)) return 1;
        return export_phantom(t, resolution, false, mask, nullptr, bvf);
    }
    return 4;
}

// `spinwalk phantom -p -i ply_file -f fov -z resolution`: mask [res][res][res] of phantom::ply::run(false).
int swref_phantom_ply(float fov_um, uint64_t resolution, const char *ply_file, uint8_t *mask, float *bvf)
{
    quiet_cout q;
    try {
        phantom::ply p(fov_um, resolution, 0.11e-6f, -1.f, ply_file, "unused.h5");
        if (!p.run(false)) return 1;
        return export_phantom(p, resolution, false, mask, nullptr, bvf);
    } catch (const std::exception &) { // happly throws on malformed files
        return 5;
    }
}

// `spinwalk dwi -b b... -v x y z -d start delta DELTA -c config` (src/spinwalk.cpp:110-114): edits `config` in place.
int swref_dwi(const double *b, uint32_t n_b, const float dir[3], uint32_t start_ms, uint32_t delta_ms, uint32_t DELTA_ms, const char *config)
{
    quiet_cout q;
    dMRI::execute_args a;
    a.start_ms = start_ms;
    a.delta_ms = delta_ms;
    a.DELTA_ms = DELTA_ms;
    a.dir = {dir[0], dir[1], dir[2]};
    a.b_value.assign(b, b + n_b);
    a.output = config;
    return dMRI::handler::execute(a) ? 0 : 1;
}

// `spinwalk config -s seq -p phantoms... -e TE -t timestep -o output` (src/spinwalk.cpp:117-121)
int swref_config(const char *seq_name, uint32_t TE_us, uint32_t timestep_us, const char *const *phantoms, uint32_t n_phantoms, const char *output)
{
    quiet_cout q;
    config::execute_args a;
    a.seq_name = seq_name;
    a.TE_us = TE_us;
    a.timestep_us = timestep_us;
    for (uint32_t i = 0; i < n_phantoms; i++) a.phantoms.push_back(phantoms[i]);
    a.output = output;
    return config::handler::execute(a) ? 0 : 1;
}

// mINI itself: read `path`, apply n edits, INIFile::write(ini, pretty) — what src/dwi/pgse.cpp:93-95,145 does around its edits.
// ops[i]: 0 ini[s][k] = v, 1 ini[s].remove(k), 2 ini.remove(s), 3 ini[s].
int swref_ini_edit(const char *path, int n, const int *ops, const char *const *sections, const char *const *keys, const char *const *values, int pretty)
{
    mINI::INIFile file(path);
    mINI::INIStructure ini;
    file.read(ini);
    for (int i = 0; i < n; i++) {
        if (ops[i] == 0) ini[sections[i]][keys[i]] = values[i];
        else if (ops[i] == 1) { if (ini.has(sections[i])) ini[sections[i]].remove(keys[i]); }
        else if (ops[i] == 2) ini.remove(sections[i]);
        else ini[sections[i]];
    }
    return file.write(ini, pretty != 0) ? 0 : 1;
}

} // extern "C"
