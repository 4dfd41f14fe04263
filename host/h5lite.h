// host/h5lite.h — a minimal HDF5 codec for exactly what `spinwalk sim` reads and writes, with no libhdf5.
//
// The reference goes through HighFive -> libhdf5 (src/sim/h5_helper.h:48-129).  libhdf5 is not available to this
// project's toolchain, so the two operations the hot path's callers need are restated from the HDF5 File Format
// Specification (version 3.0):
//   READ  numeric datasets in the ROOT group of files written by libhdf5-based tools (the reference's own `phantom`
//         subcommand via HighFive, h5py, MATLAB -v7.3): superblock v0-v3 (with or without a user block), object headers
//         v1/v2 with continuation blocks, old-style groups (symbol table: v1 B-tree + local heap + SNOD) and new-style
//         compact groups (link messages), dataspace v1/v2, fixed-point and IEEE floating-point datatypes of either byte
//         order, contiguous / compact / chunked (v1 B-tree index) layouts, deflate + shuffle + fletcher32 filters.
//         Values are converted to the requested element type like H5Dread does for HighFive's read_raw<T>
//         (h5_helper.h:69: an int8 mask becomes uint8, a float64 fieldmap becomes float32).
//   WRITE a new file holding contiguous little-endian datasets in the root group, in the layout libhdf5 produces with
//         default ("earliest") format bounds: superblock v0, v1 object headers, symbol-table root group — what the
//         reference's outputs look like (monte_carlo.cu:168-197) and what h5py / MATLAB / HighFive open.
// Out of scope (reported as errors, never silently misread): dense (fractal-heap) groups, nested groups, compound /
// variable-length / string datatypes, external and virtual storage, v2-B-tree / extensible-array chunk indexes.
#pragma once

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

namespace swk_host {
namespace h5 {

enum class DType { U8, I8, U16, I16, U32, I32, U64, I64, F32, F64 };
size_t dtype_size(DType t);
template <class T> struct dtype_of;
template <> struct dtype_of<uint8_t> { static constexpr DType value = DType::U8; };
template <> struct dtype_of<int8_t> { static constexpr DType value = DType::I8; };
template <> struct dtype_of<uint16_t> { static constexpr DType value = DType::U16; };
template <> struct dtype_of<int16_t> { static constexpr DType value = DType::I16; };
template <> struct dtype_of<uint32_t> { static constexpr DType value = DType::U32; };
template <> struct dtype_of<int32_t> { static constexpr DType value = DType::I32; };
template <> struct dtype_of<uint64_t> { static constexpr DType value = DType::U64; };
template <> struct dtype_of<int64_t> { static constexpr DType value = DType::I64; };
template <> struct dtype_of<float> { static constexpr DType value = DType::F32; };
template <> struct dtype_of<double> { static constexpr DType value = DType::F64; };

struct DatasetInfo {
    std::vector<uint64_t> dims; // row-major, slowest first (as stored)
    DType dtype = DType::U8;    // element type in the file
    bool big_endian = false;
    // storage
    int layout = 1;             // 0 compact, 1 contiguous, 2 chunked
    uint64_t address = 0, size = 0;       // contiguous: absolute file offset / bytes; chunked: B-tree address
    std::vector<uint8_t> compact;         // compact: the bytes
    std::vector<uint32_t> chunk;          // chunked: chunk dims (+ element size last)
    std::vector<std::pair<int, std::vector<uint32_t>>> filters; // (id, client data) in pipeline order
    uint64_t count() const { uint64_t n = 1; for (auto d : dims) n *= d; return n; }
};

class Reader {
public:
    ~Reader() { close(); }
    bool open(const std::string &path);
    void close();
    const std::string &error() const { return err_; }
    std::vector<std::string> names() const; // datasets / links of the root group
    bool exists(const std::string &name) const { return links_.count(name) == 1; }
    bool info(const std::string &name, DatasetInfo &out);
    // reads the whole dataset converted to `as`; dst must hold info.count() elements
    bool read(const std::string &name, DType as, void *dst, uint64_t dst_elems);
    template <class T> bool read(const std::string &name, std::vector<T> &out)
    {
        DatasetInfo di;
        if (!info(name, di)) return false;
        out.resize(di.count());
        return read(name, dtype_of<T>::value, out.data(), out.size());
    }

private:
    struct Msg { uint16_t type; uint8_t flags; std::vector<uint8_t> data; };
    bool fail(const std::string &m) { err_ = m; return false; }
    bool pread(uint64_t off, void *dst, size_t n);
    bool read_header(uint64_t addr, std::vector<Msg> &out);
    bool load_root(uint64_t root_oh, uint64_t btree, uint64_t heap);
    bool walk_group_btree(uint64_t node, uint64_t heap_data, uint64_t heap_size, int depth);
    bool read_chunked(const DatasetInfo &di, std::vector<uint8_t> &raw);
    bool walk_chunk_btree(uint64_t node, const DatasetInfo &di, std::vector<uint8_t> &raw, int depth);

    FILE *f_ = nullptr;
    uint64_t base_ = 0, file_size_ = 0;
    int size_off_ = 8, size_len_ = 8;
    std::map<std::string, uint64_t> links_; // name -> object header address (absolute)
    std::string err_;
};

// Writes all datasets of one file in one go (the reference deletes the output file and re-creates it for every run,
// monte_carlo.cu:178-179).  Data pointers must stay valid until close().
class Writer {
public:
    explicit Writer(const std::string &path) : path_(path) {}
    void add(const std::string &name, const std::vector<uint64_t> &dims, DType t, const void *data);
    template <class T> void add(const std::string &name, const std::vector<uint64_t> &dims, const std::vector<T> &v)
    {
        add(name, dims, dtype_of<T>::value, v.data());
    }
    bool close(); // writes the file; false + error() on failure
    const std::string &error() const { return err_; }

private:
    struct Item { std::string name; std::vector<uint64_t> dims; DType t; const void *data; };
    std::string path_, err_;
    std::vector<Item> items_;
};

} // namespace h5
} // namespace swk_host
