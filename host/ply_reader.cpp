#include "ply_reader.h"

#include <cstring>
#include <fstream>
#include <sstream>

namespace swk_host {

namespace {

enum class T { I8, U8, I16, U16, I32, U32, F32, F64, Bad };

T type_of(const std::string &s)
{
    if (s == "char" || s == "int8") return T::I8;
    if (s == "uchar" || s == "uint8") return T::U8;
    if (s == "short" || s == "int16") return T::I16;
    if (s == "ushort" || s == "uint16") return T::U16;
    if (s == "int" || s == "int32") return T::I32;
    if (s == "uint" || s == "uint32") return T::U32;
    if (s == "float" || s == "float32") return T::F32;
    if (s == "double" || s == "float64") return T::F64;
    return T::Bad;
}
size_t size_of(T t)
{
    switch (t) {
    case T::I8: case T::U8: return 1;
    case T::I16: case T::U16: return 2;
    case T::I32: case T::U32: case T::F32: return 4;
    case T::F64: return 8;
    default: return 0;
    }
}
bool is_float(T t) { return t == T::F32 || t == T::F64; }

struct Property { std::string name; bool list = false; T count_type = T::Bad, type = T::Bad; };
struct Element { std::string name; size_t count = 0; std::vector<Property> props; };

// one scalar of a binary body as double (exact for every PLY type)
bool read_binary(std::istream &in, T t, bool swap, double &out)
{
    unsigned char b[8];
    const size_t n = size_of(t);
    if (!in.read(reinterpret_cast<char *>(b), std::streamsize(n))) return false;
    if (swap)
        for (size_t i = 0; i < n / 2; i++) std::swap(b[i], b[n - 1 - i]);
    switch (t) {
    case T::I8: { int8_t v; memcpy(&v, b, 1); out = v; break; }
    case T::U8: { uint8_t v; memcpy(&v, b, 1); out = v; break; }
    case T::I16: { int16_t v; memcpy(&v, b, 2); out = v; break; }
    case T::U16: { uint16_t v; memcpy(&v, b, 2); out = v; break; }
    case T::I32: { int32_t v; memcpy(&v, b, 4); out = v; break; }
    case T::U32: { uint32_t v; memcpy(&v, b, 4); out = v; break; }
    case T::F32: { float v; memcpy(&v, b, 4); out = v; break; }
    case T::F64: { double v; memcpy(&v, b, 8); out = v; break; }
    default: return false;
    }
    return true;
}

// one scalar of an ascii body; float properties are read at their own precision (a `float` token is rounded to float first)
bool read_ascii(std::istream &in, T t, double &out)
{
    std::string tok;
    if (!(in >> tok)) return false;
    try {
        size_t used = 0;
        if (t == T::F32) out = std::stof(tok, &used);
        else if (t == T::F64) out = std::stod(tok, &used);
        else out = double(std::stoll(tok, &used));
        return used == tok.size();
    } catch (const std::exception &) {
        return false;
    }
}

bool host_is_little_endian()
{
    const uint16_t x = 1;
    unsigned char b;
    memcpy(&b, &x, 1);
    return b == 1;
}

} // namespace

bool read_ply(const std::string &path, PlyMesh &mesh, std::string &error)
{
    mesh.vertices.clear();
    mesh.faces.clear();
    std::ifstream in(path, std::ios::in | std::ios::binary);
    if (!in.is_open()) { error = "could not open ply file: " + path; return false; }

    // ---- header
    auto getline = [&](std::string &line) {
        if (!std::getline(in, line)) return false;
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        return true;
    };
    std::string line;
    if (!getline(line) || line != "ply") { error = "not a ply file (no \"ply\" magic): " + path; return false; }
    std::string format;
    std::vector<Element> elements;
    bool ended = false;
    while (getline(line)) {
        std::istringstream ls(line);
        std::string word;
        if (!(ls >> word)) continue;
        if (word == "comment" || word == "obj_info") continue;
        if (word == "format") {
            std::string version;
            ls >> format >> version;
            if (format != "ascii" && format != "binary_little_endian" && format != "binary_big_endian") { error = "unknown ply format: " + format; return false; }
        } else if (word == "element") {
            Element e;
            if (!(ls >> e.name >> e.count)) { error = "malformed element line: " + line; return false; }
            elements.push_back(e);
        } else if (word == "property") {
            if (elements.empty()) { error = "property before any element: " + line; return false; }
            Property p;
            std::string a, b, c;
            ls >> a;
            if (a == "list") {
                p.list = true;
                if (!(ls >> b >> c >> p.name)) { error = "malformed list property: " + line; return false; }
                p.count_type = type_of(b);
                p.type = type_of(c);
                if (p.count_type == T::Bad || is_float(p.count_type)) { error = "bad list count type: " + line; return false; }
            } else {
                p.type = type_of(a);
                if (!(ls >> p.name)) { error = "malformed property: " + line; return false; }
            }
            if (p.type == T::Bad) { error = "unknown property type: " + line; return false; }
            elements.back().props.push_back(p);
        } else if (word == "end_header") { ended = true; break; }
        else { error = "unrecognised header line: " + line; return false; }
    }
    if (!ended || format.empty()) { error = "incomplete ply header: " + path; return false; }

    const bool ascii = format == "ascii";
    const bool swap = !ascii && ((format == "binary_little_endian") != host_is_little_endian());
    bool have_vertices = false, have_faces = false;

    // ---- body, element by element in file order
    for (const Element &e : elements) {
        const bool is_vertex = e.name == "vertex", is_face = e.name == "face";
        int ix = -1, iy = -1, iz = -1, il = -1;
        for (size_t k = 0; k < e.props.size(); k++) {
            const Property &p = e.props[k];
            if (is_vertex && !p.list && p.name == "x") ix = int(k);
            if (is_vertex && !p.list && p.name == "y") iy = int(k);
            if (is_vertex && !p.list && p.name == "z") iz = int(k);
            if (is_face && p.list && (p.name == "vertex_indices" || p.name == "vertex_index") && il < 0) il = int(k);
        }
        if (is_vertex) {
            if (ix < 0 || iy < 0 || iz < 0) { error = "ply element \"vertex\" has no x / y / z properties"; return false; }
            for (int k : {ix, iy, iz})
                if (!is_float(e.props[k].type)) { error = "vertex positions must be float or double properties"; return false; }
            mesh.vertices.resize(e.count * 3);
            have_vertices = true;
        }
        if (is_face) {
            if (il < 0) { error = "ply element \"face\" has no vertex_indices / vertex_index list"; return false; }
            if (is_float(e.props[il].type)) { error = "face indices must be an integer list"; return false; }
            mesh.faces.resize(e.count * 3);
            have_faces = true;
        }
        for (size_t item = 0; item < e.count; item++) {
            std::istringstream ls;
            if (ascii) { // one item per line
                do {
                    if (!std::getline(in, line)) { error = "ply body ends early in element " + e.name; return false; }
                } while (line.find_first_not_of(" \t\r") == std::string::npos);
                ls.str(line);
            }
            std::istream &src = ascii ? static_cast<std::istream &>(ls) : static_cast<std::istream &>(in);
            for (size_t k = 0; k < e.props.size(); k++) {
                const Property &p = e.props[k];
                double v = 0;
                if (!p.list) {
                    if (!(ascii ? read_ascii(src, p.type, v) : read_binary(src, p.type, swap, v))) { error = "ply body is malformed or ends early in element " + e.name; return false; }
                    if (is_vertex && int(k) == ix) mesh.vertices[3 * item + 0] = v;
                    if (is_vertex && int(k) == iy) mesh.vertices[3 * item + 1] = v;
                    if (is_vertex && int(k) == iz) mesh.vertices[3 * item + 2] = v;
                    continue;
                }
                double cnt = 0;
                if (!(ascii ? read_ascii(src, p.count_type, cnt) : read_binary(src, p.count_type, swap, cnt)) || cnt < 0) { error = "ply body is malformed or ends early in element " + e.name; return false; }
                const size_t n = size_t(cnt);
                if (is_face && int(k) == il && n != 3) { error = "Only triangular mesh is supported!"; return false; } // phantom_ply.cpp:150-154
                for (size_t j = 0; j < n; j++) {
                    if (!(ascii ? read_ascii(src, p.type, v) : read_binary(src, p.type, swap, v))) { error = "ply body is malformed or ends early in element " + e.name; return false; }
                    if (is_face && int(k) == il) {
                        if (v < 0) { error = "negative vertex index in a face"; return false; }
                        mesh.faces[3 * item + j] = uint64_t(v);
                    }
                }
            }
        }
    }
    if (!have_vertices) { error = "ply file has no \"vertex\" element"; return false; }
    if (!have_faces) { error = "ply file has no \"face\" element"; return false; }
    for (uint64_t i : mesh.faces)
        if (i >= mesh.n_vertices()) { error = "face index out of range"; return false; }
    return true;
}

} // namespace swk_host
