// host/ply_reader.h — reads the triangle mesh of a PLY file for `spinwalk phantom -p -i mesh.ply`.
//
// The reference uses the vendored happly (include/happly.h) and takes two things from the file (src/phantom/phantom_ply.cpp:146-156):
// getVertexPositions() — the x, y, z properties of element "vertex" as double — and getFaceIndices<size_t>() — the list property
// "vertex_indices" (or "vertex_index") of element "face"; faces that are not triangles are an error.  This reader restates the PLY
// 1.0 format for exactly that: ascii, binary_little_endian and binary_big_endian bodies, every scalar type name of the format
// (char/int8 ... double/float64), comment / obj_info lines, and elements or properties it does not need are parsed and skipped.
#pragma once

#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace swk_host {

struct PlyMesh {
    std::vector<double> vertices;   // [n][3]
    std::vector<uint64_t> faces;    // [m][3]
    size_t n_vertices() const { return vertices.size() / 3; }
    size_t n_faces() const { return faces.size() / 3; }
};

// false + error on malformed files, missing x/y/z or face lists, non-float positions, negative or non-triangular faces
bool read_ply(const std::string &path, PlyMesh &mesh, std::string &error);

} // namespace swk_host
