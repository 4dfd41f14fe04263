// host/phantom_cli.cpp — `spinwalk phantom` (-c / -s / -t / -p): the reference's phantom::handler::execute (src/phantom/handler.cpp:10-35) on the GPU
// generator (include/spinwalk_phantom.h), writing the phantom file of phantom_base::save (src/phantom/phantom_base.cpp:69-103):
// datasets /fieldmap (float32 [n,n,n], only when oxy_level >= 0), /mask (uint8 [n,n,n]), /fov (float32 [3], metres), /bvf (float32 [1]).
#include <cstdio>
#include <filesystem>
#include <vector>

#include "../include/spinwalk_phantom.h"
#include "generators.h"
#include "h5lite.h"
#include "ply_reader.h"

namespace fs = std::filesystem;

namespace swk_host {

namespace {
bool one_phantom(const PhantomArgs &a, int shape, const char *what, std::string &error)
{
    if (!a.quiet) printf("Generating %s...\n", what);
    swk_phantom_spec sp{};
    sp.shape = shape;
    sp.fov_um = a.fov;
    sp.resolution = a.resolution;
    sp.radius_um = a.radius;
    sp.volume_fraction = a.volume_fraction;
    sp.orientation_deg = a.orientation;
    sp.seed = a.seed;
    sp.dchi = a.dchi;
    sp.oxy_level = a.oxy_level;
    if (shape == SWK_SHAPE_TWOPOOLS) { // twopools(fov, resolution, filename) -> phantom_base(fov, res, 0, -1, 0, 0, filename) (phantom_twopools.cpp:29-30)
        sp.dchi = 0.f;
        sp.oxy_level = -1.f;
        sp.volume_fraction = 0.f;
        sp.seed = 0;
    }
    const size_t n = a.resolution, V = n * n * n;
    const bool field = shape != SWK_SHAPE_TWOPOOLS && sp.oxy_level >= 0;
    std::vector<uint8_t> mask(V);
    std::vector<float> fieldmap(field ? V : 0);
    swk_phantom_stats st{};
    if (swk_phantom_generate(a.device, &sp, mask.data(), field ? fieldmap.data() : nullptr, 0, &st) != SWK_OK) {
        error = swk_phantom_last_error();
        return false;
    }
    if (!a.quiet)
        printf("%u shapes, actual volume fraction = %g %%, placement %.1f ms, voxel fill on the GPU %.2f ms\n", st.n_shapes, st.volume_fraction, st.place_ms,
               st.kernel_ms);

    // phantom_base::save
    const fs::path parent = fs::absolute(a.output).parent_path();
    std::error_code ec;
    if (!fs::is_directory(parent) && !fs::create_directories(parent, ec)) { error = "cannot create directory " + parent.string(); return false; }
    const float fov_m[3] = {a.fov * 1e-6f, a.fov * 1e-6f, a.fov * 1e-6f};
    h5::Writer w(a.output);
    if (field) w.add("fieldmap", {n, n, n}, h5::DType::F32, fieldmap.data());
    w.add("mask", {n, n, n}, h5::DType::U8, mask.data());
    w.add("fov", {3}, h5::DType::F32, fov_m);
    w.add("bvf", {1}, h5::DType::F32, &st.volume_fraction);
    if (!w.close()) { error = w.error(); return false; }
    return true;
}
// ≙ phantom::ply::run(true) (src/phantom/phantom_ply.cpp:141-255): mask only; the reference leaves `bvf` at 0 for mesh phantoms (:187)
bool mesh_phantom(const PhantomArgs &a, std::string &error)
{
    if (!a.quiet) printf("Generating phantom from triangular mesh...\n");
    PlyMesh mesh;
    if (!read_ply(a.ply_file, mesh, error)) return false;
    const size_t n = a.resolution, V = n * n * n;
    std::vector<uint8_t> mask(V);
    swk_phantom_stats st{};
    if (swk_phantom_mesh(a.device, a.fov, a.resolution, mesh.vertices.data(), mesh.n_vertices(), mesh.faces.data(), mesh.n_faces(), mask.data(), 0, &st) != SWK_OK) {
        error = swk_phantom_last_error();
        return false;
    }
    if (!a.quiet)
        printf("%zu vertices, %zu triangles, %g %% of the voxels inside; mesh preparation %.1f ms, voxel fill on the GPU %.2f ms\n", mesh.n_vertices(), mesh.n_faces(),
               st.volume_fraction, st.place_ms, st.kernel_ms);
    const fs::path parent = fs::absolute(a.output).parent_path();
    std::error_code ec;
    if (!fs::is_directory(parent) && !fs::create_directories(parent, ec)) { error = "cannot create directory " + parent.string(); return false; }
    const float fov_m[3] = {a.fov * 1e-6f, a.fov * 1e-6f, a.fov * 1e-6f};
    const float bvf = 0.f;
    h5::Writer w(a.output);
    w.add("mask", {n, n, n}, h5::DType::U8, mask.data());
    w.add("fov", {3}, h5::DType::F32, fov_m);
    w.add("bvf", {1}, h5::DType::F32, &bvf);
    if (!w.close()) { error = w.error(); return false; }
    return true;
}
} // namespace

bool generate_phantom(const PhantomArgs &a, std::string &error)
{
    bool ok = true; // like the reference, every selected shape is generated in turn into the same output file
    if (a.cylinder) ok = ok && one_phantom(a, SWK_SHAPE_CYLINDER, "cylinder phantom", error);
    if (a.sphere) ok = ok && one_phantom(a, SWK_SHAPE_SPHERE, "sphere phantom", error);
    if (a.twopools) ok = ok && one_phantom(a, SWK_SHAPE_TWOPOOLS, "phantom with two pools", error);
    if (a.ply) ok = ok && mesh_phantom(a, error);
    if (ok && !a.quiet) printf("Done.\n");
    return ok;
}

} // namespace swk_host
