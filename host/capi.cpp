// host/capi.cpp — C hooks into the host-side pieces (config reader, HDF5 codec) for the Python tests (ctypes).
#include <cstring>
#include <string>

#include "sim_config.h"

extern "C" {

// Parses `path` like `spinwalk sim -c path` would and writes the resulting configuration as JSON into buf.
// Returns 0 on success, 1 when the config is rejected (buf then holds {"ok": false, "error": "..."}), -1 if buf is too small.
int swkh_config_json(const char *path, int check_files, char *buf, size_t n)
{
    swk_host::SimConfig c;
    std::string out;
    int rc = 0;
    if (c.prepare(path, check_files != 0)) {
        out = c.to_json();
        out.insert(1, "\"ok\": true, ");
    } else {
        std::string e;
        for (char ch : c.error) {
            if (ch == '"' || ch == '\\') e += '\\';
            e += ch;
        }
        out = "{\"ok\": false, \"error\": \"" + e + "\"}";
        rc = 1;
    }
    if (out.size() + 1 > n) return -1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return rc;
}

} // extern "C"

#include "h5lite.h"

namespace {
thread_local std::string g_h5_error;
int copy_out(const std::string &s, char *buf, size_t n)
{
    if (s.size() + 1 > n) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}
} // namespace

extern "C" {

const char *swkh_h5_error(void) { return g_h5_error.c_str(); }

// names of the root group's links, '\n'-separated
int swkh_h5_names(const char *path, char *buf, size_t n)
{
    swk_host::h5::Reader r;
    if (!r.open(path)) { g_h5_error = r.error(); return 1; }
    std::string s;
    for (const auto &nm : r.names()) s += nm + "\n";
    return copy_out(s, buf, n);
}

// dtype codes = swk_host::h5::DType order: u8 i8 u16 i16 u32 i32 u64 i64 f32 f64
int swkh_h5_info(const char *path, const char *name, int *rank, uint64_t *dims /*[8]*/, int *dtype, int *layout)
{
    swk_host::h5::Reader r;
    swk_host::h5::DatasetInfo di;
    if (!r.open(path) || !r.info(name, di)) { g_h5_error = r.error(); return 1; }
    if (di.dims.size() > 8) { g_h5_error = "rank > 8"; return 1; }
    *rank = (int)di.dims.size();
    for (size_t i = 0; i < di.dims.size(); i++) dims[i] = di.dims[i];
    *dtype = (int)di.dtype;
    if (layout) *layout = di.layout;
    return 0;
}

int swkh_h5_read(const char *path, const char *name, int as_dtype, void *dst, uint64_t n_elems)
{
    swk_host::h5::Reader r;
    if (!r.open(path) || !r.read(name, (swk_host::h5::DType)as_dtype, dst, n_elems)) { g_h5_error = r.error(); return 1; }
    return 0;
}

int swkh_h5_write(const char *path, int n, const char *const *names, const int *ranks, const uint64_t *dims_flat, const int *dtypes, const void *const *data)
{
    swk_host::h5::Writer w(path);
    size_t q = 0;
    for (int i = 0; i < n; i++) {
        std::vector<uint64_t> d(dims_flat + q, dims_flat + q + ranks[i]);
        q += ranks[i];
        w.add(names[i], d, (swk_host::h5::DType)dtypes[i], data[i]);
    }
    if (!w.close()) { g_h5_error = w.error(); return 1; }
    return 0;
}

} // extern "C"

#include "generators.h"
#include "ini_edit.h"

extern "C" {

// `spinwalk dwi` (host/generators.cpp): edits `config` in place.  0 ok, 1 failed (message in buf).
int swkh_dwi(const double *b, uint32_t n_b, const float dir[3], uint32_t start_ms, uint32_t delta_ms, uint32_t DELTA_ms, const char *config, char *buf, size_t n)
{
    swk_host::DwiArgs a;
    a.b_value.assign(b, b + n_b);
    a.dir = {dir[0], dir[1], dir[2]};
    a.start_ms = start_ms;
    a.delta_ms = delta_ms;
    a.DELTA_ms = DELTA_ms;
    a.config = config;
    std::string err;
    const bool ok = swk_host::generate_dwi(a, err);
    if (buf && n) copy_out(err, buf, n);
    return ok ? 0 : 1;
}

// `spinwalk config`
int swkh_config(const char *seq_name, uint32_t TE_us, uint32_t timestep_us, const char *const *phantoms, uint32_t n_phantoms, const char *output, char *buf, size_t n)
{
    swk_host::ConfigArgs a;
    a.seq_name = seq_name;
    a.TE_us = TE_us;
    a.timestep_us = timestep_us;
    for (uint32_t i = 0; i < n_phantoms; i++) a.phantoms.push_back(phantoms[i]);
    a.output = output;
    std::string err;
    const bool ok = swk_host::generate_config(a, err);
    if (buf && n) copy_out(err, buf, n);
    return ok ? 0 : 1;
}

// IniDocument round trip for the writer tests: parse `path`, apply n edits, update the file in place (pretty or not).
// ops[i]: 0 set(section,key,value), 1 remove key, 2 remove section, 3 touch section.
int swkh_ini_edit(const char *path, int n, const int *ops, const char *const *sections, const char *const *keys, const char *const *values, int pretty)
{
    swk_host::IniDocument d;
    d.load(path); // a missing file gives an empty document, update_file then creates it
    for (int i = 0; i < n; i++) {
        if (ops[i] == 0) d.set(sections[i], keys[i], values[i]);
        else if (ops[i] == 1) d.remove(sections[i], keys[i]);
        else if (ops[i] == 2) d.remove_section(sections[i]);
        else d.touch_section(sections[i]);
    }
    return d.update_file(path, pretty != 0) ? 0 : 1;
}

} // extern "C"

#include "ply_reader.h"

extern "C" {

// PLY reader for the tests: call with vertices == NULL to get the sizes, again with buffers to get the data.  0 ok, 1 failed (message in buf).
int swkh_ply_read(const char *path, double *vertices, uint64_t *n_vertices, uint64_t *faces, uint64_t *n_faces, char *buf, size_t n)
{
    swk_host::PlyMesh m;
    std::string err;
    if (!swk_host::read_ply(path, m, err)) {
        if (buf && n) copy_out(err, buf, n);
        return 1;
    }
    if (vertices && *n_vertices >= m.n_vertices()) memcpy(vertices, m.vertices.data(), m.vertices.size() * sizeof(double));
    if (faces && *n_faces >= m.n_faces()) memcpy(faces, m.faces.data(), m.faces.size() * sizeof(uint64_t));
    *n_vertices = m.n_vertices();
    *n_faces = m.n_faces();
    return 0;
}

} // extern "C"
