// host/h5lite.cpp — see h5lite.h.  Field layouts follow the HDF5 File Format Specification 3.0 (sections II.A superblock,
// III.A B-trees, III.B symbol table nodes, III.D local heaps, IV.A object headers and messages).
#include "h5lite.h"

#include <algorithm>
#include <cstring>

#include <zlib.h>

namespace swk_host {
namespace h5 {

namespace {

const uint8_t kSig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
const uint64_t kUndef = ~0ull;

uint64_t rdn(const uint8_t *p, int n)
{
    uint64_t v = 0;
    for (int i = 0; i < n && i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    if (n < 8 && v == ((1ull << (8 * n)) - 1)) return kUndef; // undefined address of a narrower width
    return v;
}
uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

void put(std::vector<uint8_t> &b, uint64_t v, int n)
{
    for (int i = 0; i < n; i++) b.push_back(i < 8 ? (uint8_t)(v >> (8 * i)) : (uint8_t)0); // n > 8: zero fill
}
void pad8(std::vector<uint8_t> &b)
{
    while (b.size() % 8) b.push_back(0);
}

template <class S, class D>
void conv_loop(const uint8_t *src, void *dst, uint64_t n, bool swap)
{
    D *d = static_cast<D *>(dst);
    for (uint64_t i = 0; i < n; i++) {
        S v;
        uint8_t t[sizeof(S)];
        if (swap) {
            for (size_t k = 0; k < sizeof(S); k++) t[k] = src[i * sizeof(S) + sizeof(S) - 1 - k];
            memcpy(&v, t, sizeof(S));
        } else {
            memcpy(&v, src + i * sizeof(S), sizeof(S));
        }
        d[i] = static_cast<D>(v);
    }
}
template <class S>
void conv_from(const uint8_t *src, DType to, void *dst, uint64_t n, bool swap)
{
    switch (to) {
    case DType::U8: conv_loop<S, uint8_t>(src, dst, n, swap); break;
    case DType::I8: conv_loop<S, int8_t>(src, dst, n, swap); break;
    case DType::U16: conv_loop<S, uint16_t>(src, dst, n, swap); break;
    case DType::I16: conv_loop<S, int16_t>(src, dst, n, swap); break;
    case DType::U32: conv_loop<S, uint32_t>(src, dst, n, swap); break;
    case DType::I32: conv_loop<S, int32_t>(src, dst, n, swap); break;
    case DType::U64: conv_loop<S, uint64_t>(src, dst, n, swap); break;
    case DType::I64: conv_loop<S, int64_t>(src, dst, n, swap); break;
    case DType::F32: conv_loop<S, float>(src, dst, n, swap); break;
    case DType::F64: conv_loop<S, double>(src, dst, n, swap); break;
    }
}
void convert(const uint8_t *src, DType from, bool swap, DType to, void *dst, uint64_t n)
{
    switch (from) {
    case DType::U8: conv_from<uint8_t>(src, to, dst, n, swap); break;
    case DType::I8: conv_from<int8_t>(src, to, dst, n, swap); break;
    case DType::U16: conv_from<uint16_t>(src, to, dst, n, swap); break;
    case DType::I16: conv_from<int16_t>(src, to, dst, n, swap); break;
    case DType::U32: conv_from<uint32_t>(src, to, dst, n, swap); break;
    case DType::I32: conv_from<int32_t>(src, to, dst, n, swap); break;
    case DType::U64: conv_from<uint64_t>(src, to, dst, n, swap); break;
    case DType::I64: conv_from<int64_t>(src, to, dst, n, swap); break;
    case DType::F32: conv_from<float>(src, to, dst, n, swap); break;
    case DType::F64: conv_from<double>(src, to, dst, n, swap); break;
    }
}

// datatype message -> (DType, big endian).  Enumerations (h5py bool) resolve to their integer base type.
bool parse_datatype(const uint8_t *p, size_t n, DType &t, bool &be, std::string &why)
{
    if (n < 8) { why = "truncated datatype message"; return false; }
    const int cls = p[0] & 0x0f;
    const uint32_t size = rd32(p + 4);
    be = (p[1] & 1) != 0;
    if (cls == 0) { // fixed point
        const bool sgn = (p[1] & 0x08) != 0;
        switch (size) {
        case 1: t = sgn ? DType::I8 : DType::U8; return true;
        case 2: t = sgn ? DType::I16 : DType::U16; return true;
        case 4: t = sgn ? DType::I32 : DType::U32; return true;
        case 8: t = sgn ? DType::I64 : DType::U64; return true;
        }
        why = "integer datatype of " + std::to_string(size) + " bytes";
        return false;
    }
    if (cls == 1) { // floating point: IEEE binary32 / binary64 only
        if (n < 20) { why = "truncated floating-point datatype"; return false; }
        const int exp_size = p[13], man_size = p[15];
        if (size == 4 && exp_size == 8 && man_size == 23) { t = DType::F32; return true; }
        if (size == 8 && exp_size == 11 && man_size == 52) { t = DType::F64; return true; }
        why = "non-IEEE floating-point datatype (" + std::to_string(size) + " bytes)";
        return false;
    }
    if (cls == 8 && n >= 16) return parse_datatype(p + 8, n - 8, t, be, why); // enumeration: base type follows the header
    why = "datatype class " + std::to_string(cls) + " (only integers and IEEE floats are supported)";
    return false;
}

} // namespace

size_t dtype_size(DType t)
{
    switch (t) {
    case DType::U8: case DType::I8: return 1;
    case DType::U16: case DType::I16: return 2;
    case DType::U32: case DType::I32: case DType::F32: return 4;
    default: return 8;
    }
}

// =====================================================================================================================
// Reader
// =====================================================================================================================
void Reader::close()
{
    if (f_) fclose(f_);
    f_ = nullptr;
    links_.clear();
}

bool Reader::pread(uint64_t off, void *dst, size_t n)
{
    if (off == kUndef || off + n > file_size_) return fail("read past the end of the file (truncated or unsupported file)");
    if (fseeko(f_, (off_t)off, SEEK_SET) != 0) return fail("seek failed");
    if (n && fread(dst, 1, n, f_) != n) return fail("read failed");
    return true;
}

bool Reader::open(const std::string &path)
{
    close();
    err_.clear();
    f_ = fopen(path.c_str(), "rb");
    if (!f_) return fail("cannot open " + path);
    fseeko(f_, 0, SEEK_END);
    file_size_ = (uint64_t)ftello(f_);
    // the superblock sits at 0 or, behind a user block (MATLAB -v7.3: 512 bytes), at 512, 1024, 2048, ...
    uint64_t sb = kUndef;
    uint8_t h[128];
    for (uint64_t off = 0; off + 8 <= file_size_; off = off ? off * 2 : 512) {
        uint8_t s[8];
        if (!pread(off, s, 8)) break;
        if (memcmp(s, kSig, 8) == 0) { sb = off; break; }
    }
    if (sb == kUndef) return fail(path + " is not an HDF5 file (no superblock signature)");
    const size_t hn = (size_t)std::min<uint64_t>(sizeof h, file_size_ - sb);
    memset(h, 0, sizeof h);
    if (!pread(sb, h, hn)) return false;
    const int ver = h[8];
    uint64_t root_oh = kUndef, btree = kUndef, heap = kUndef, base_field = 0;
    auto sizes_ok = [&] { return (size_off_ == 2 || size_off_ == 4 || size_off_ == 8) && (size_len_ == 2 || size_len_ == 4 || size_len_ == 8); };
    if (ver == 0 || ver == 1) {
        size_off_ = h[13];
        size_len_ = h[14];
        if (!sizes_ok()) return fail("unsupported size of offsets / lengths"); // they are strides into h[] below
        size_t p = (ver == 0) ? 24 : 28;
        base_field = rdn(h + p, size_off_);
        p += 4 * (size_t)size_off_; // base, free-space info, end of file, driver info
        // root group symbol table entry: link name offset, object header address, cache type, reserved, scratch pad
        root_oh = rdn(h + p + size_off_, size_off_);
        const uint32_t cache = rd32(h + p + 2 * size_off_);
        if (cache == 1) {
            btree = rdn(h + p + 2 * size_off_ + 8, size_off_);
            heap = rdn(h + p + 3 * size_off_ + 8, size_off_);
        }
    } else if (ver == 2 || ver == 3) {
        size_off_ = h[9];
        size_len_ = h[10];
        if (!sizes_ok()) return fail("unsupported size of offsets / lengths");
        base_field = rdn(h + 12, size_off_);
        root_oh = rdn(h + 12 + 3 * size_off_, size_off_);
    } else {
        return fail("superblock version " + std::to_string(ver) + " is not supported");
    }
    base_ = (base_field && base_field != kUndef) ? base_field : sb; // addresses are relative to the base address (= user block size)
    if (root_oh == kUndef) return fail("root group has no object header");
    return load_root(base_ + root_oh, btree, heap);
}

// Object header (v1 or v2) -> flat list of messages, continuation blocks followed.
bool Reader::read_header(uint64_t addr, std::vector<Msg> &out)
{
    out.clear();
    uint8_t p[64] = {0};
    if (!pread(addr, p, (size_t)std::min<uint64_t>(sizeof p, file_size_ - std::min(addr, file_size_)))) return false;
    std::vector<std::pair<uint64_t, uint64_t>> chunks; // (absolute offset of message area, length)
    const bool v2 = memcmp(p, "OHDR", 4) == 0;
    uint32_t n_msgs_v1 = 0;
    uint8_t oh_flags = 0;
    if (v2) {
        if (p[4] != 2) return fail("object header version " + std::to_string(p[4]) + " is not supported");
        oh_flags = p[5];
        size_t q = 6;
        if (oh_flags & 0x20) q += 16; // access, modification, change, birth times
        if (oh_flags & 0x10) q += 4;  // max compact / min dense attributes
        const int w = 1 << (oh_flags & 3);
        const uint64_t size0 = rdn(p + q, w) == kUndef ? 0 : rdn(p + q, w);
        chunks.push_back({addr + q + w, size0});
    } else {
        if (p[0] != 1) return fail("not an object header (version byte " + std::to_string(p[0]) + ")");
        n_msgs_v1 = rd16(p + 2);
        chunks.push_back({addr + 16, rd32(p + 8)}); // 12-byte prefix padded to 8-byte alignment
    }
    for (size_t c = 0; c < chunks.size(); c++) {
        if (chunks[c].second > (64u << 20)) return fail("implausible object header chunk size");
        std::vector<uint8_t> b((size_t)chunks[c].second);
        if (!pread(chunks[c].first, b.data(), b.size())) return false;
        size_t q = 0, end = b.size();
        if (v2 && c > 0) { // continuation chunk: "OCHK" ... checksum
            if (b.size() < 8 || memcmp(b.data(), "OCHK", 4) != 0) return fail("bad object header continuation block");
            q = 4;
            end = b.size() - 4;
        }
        const size_t mh = v2 ? (size_t)(4 + ((oh_flags & 0x04) ? 2 : 0)) : 8;
        while (q + mh <= end) {
            if (!v2 && out.size() >= n_msgs_v1 && n_msgs_v1) break;
            Msg m;
            size_t sz;
            if (v2) {
                m.type = b[q];
                sz = rd16(&b[q + 1]);
                m.flags = b[q + 3];
            } else {
                m.type = rd16(&b[q]);
                sz = rd16(&b[q + 2]);
                m.flags = b[q + 4];
            }
            q += mh;
            if (q + sz > end) break; // gap / padding at the end of a chunk
            m.data.assign(b.begin() + q, b.begin() + q + sz);
            q += sz;
            if (m.type == 0x0010) { // continuation: offset, length
                if (m.data.size() < (size_t)(size_off_ + size_len_)) return fail("bad continuation message");
                const uint64_t off = rdn(m.data.data(), size_off_), len = rdn(m.data.data() + size_off_, size_len_);
                if (off != kUndef) chunks.push_back({base_ + off, len});
                if (!v2) out.push_back(m); // counts towards the v1 message total
                continue;
            }
            out.push_back(m);
        }
        if (chunks.size() > 1024) return fail("object header continuation loop");
    }
    return true;
}

bool Reader::load_root(uint64_t root_oh, uint64_t btree, uint64_t heap)
{
    std::vector<Msg> msgs;
    if (!read_header(root_oh, msgs)) return false;
    bool dense = false;
    for (const Msg &m : msgs) {
        const uint8_t *d = m.data.data();
        if (m.type == 0x0011 && m.data.size() >= (size_t)2 * size_off_) { // symbol table: B-tree + local heap
            btree = rdn(d, size_off_);
            heap = rdn(d + size_off_, size_off_);
        } else if (m.type == 0x0006 && m.data.size() >= 4) { // link message (compact new-style group)
            size_t q = 0;
            if (d[q++] != 1) return fail("link message version");
            const uint8_t fl = d[q++];
            uint8_t type = 0;
            if (fl & 0x08) type = d[q++];
            if (fl & 0x04) q += 8;
            if (fl & 0x10) q += 1;
            const int w = 1 << (fl & 3);
            if (q + w > m.data.size()) return fail("bad link message");
            const uint64_t len = rdn(d + q, w);
            q += w;
            if (q + len + (type == 0 ? (size_t)size_off_ : 0) > m.data.size()) return fail("bad link message");
            const std::string name((const char *)d + q, (size_t)len);
            q += len;
            if (type == 0) links_[name] = base_ + rdn(d + q, size_off_); // hard link; soft / external links are skipped
        } else if (m.type == 0x0002 && m.data.size() >= (size_t)2 + size_off_) { // link info: is the group stored densely?
            size_t q = 2;
            if (d[1] & 1) q += 8;
            if (q + size_off_ <= m.data.size() && rdn(d + q, size_off_) != kUndef) dense = true;
        }
    }
    if (btree != kUndef && heap != kUndef) {
        uint8_t hh[8 + 3 * 8];
        if (!pread(base_ + heap, hh, 8 + 2 * size_len_ + size_off_)) return false;
        if (memcmp(hh, "HEAP", 4) != 0) return fail("bad local heap signature");
        const uint64_t data_size = rdn(hh + 8, size_len_), data_addr = rdn(hh + 8 + 2 * size_len_, size_off_);
        if (!walk_group_btree(base_ + btree, base_ + data_addr, data_size, 0)) return false;
    } else if (links_.empty() && dense) {
        return fail("the root group uses dense (fractal heap) link storage, which this reader does not support");
    }
    return true;
}

bool Reader::walk_group_btree(uint64_t node, uint64_t heap_data, uint64_t heap_size, int depth)
{
    if (depth > 16) return fail("group B-tree too deep");
    uint8_t h[8 + 16];
    if (!pread(node, h, 8 + 2 * size_off_)) return false;
    if (memcmp(h, "TREE", 4) != 0 || h[4] != 0) return fail("bad group B-tree node");
    const int level = h[5], n = rd16(h + 6);
    std::vector<uint8_t> b((size_t)n * (size_len_ + size_off_) + size_len_);
    if (!pread(node + 8 + 2 * size_off_, b.data(), b.size())) return false;
    for (int i = 0; i < n; i++) {
        const uint64_t child = base_ + rdn(&b[(size_t)i * (size_len_ + size_off_) + size_len_], size_off_);
        if (level > 0) {
            if (!walk_group_btree(child, heap_data, heap_size, depth + 1)) return false;
            continue;
        }
        uint8_t s[8];
        if (!pread(child, s, 8)) return false;
        if (memcmp(s, "SNOD", 4) != 0) return fail("bad symbol table node");
        const int nsym = rd16(s + 6);
        const size_t esz = 2 * (size_t)size_off_ + 8 + 16;
        std::vector<uint8_t> e(esz * nsym);
        if (!pread(child + 8, e.data(), e.size())) return false;
        for (int k = 0; k < nsym; k++) {
            const uint64_t name_off = rdn(&e[k * esz], size_off_), oh = rdn(&e[k * esz + size_off_], size_off_);
            if (name_off >= heap_size) return fail("symbol name outside the local heap");
            std::string name;
            char c[64];
            for (uint64_t o = name_off; o < heap_size;) { // NUL-terminated string in the heap data segment
                const size_t take = (size_t)std::min<uint64_t>(sizeof c, heap_size - o);
                if (!pread(heap_data + o, c, take)) return false;
                const void *z = memchr(c, 0, take);
                if (z) { name.append(c, (const char *)z - c); break; }
                name.append(c, take);
                o += take;
            }
            links_[name] = base_ + oh;
        }
    }
    return true;
}

std::vector<std::string> Reader::names() const
{
    std::vector<std::string> v;
    for (const auto &kv : links_) v.push_back(kv.first);
    return v;
}

bool Reader::info(const std::string &name, DatasetInfo &di)
{
    di = DatasetInfo();
    auto it = links_.find(name);
    if (it == links_.end()) return fail("dataset \"" + name + "\" does not exist");
    std::vector<Msg> msgs;
    if (!read_header(it->second, msgs)) return false;
    bool have_space = false, have_type = false, have_layout = false;
    for (const Msg &m : msgs) {
        const uint8_t *d = m.data.data();
        const size_t n = m.data.size();
        if (m.type == 0x0001 && n >= 4) { // dataspace
            const int ver = d[0], rank = d[1];
            const size_t q = ver == 1 ? 8 : 4;
            if (ver != 1 && ver != 2) return fail("dataspace version " + std::to_string(ver));
            if (q + (size_t)rank * size_len_ > n) return fail("bad dataspace message");
            for (int i = 0; i < rank; i++) di.dims.push_back(rdn(d + q + (size_t)i * size_len_, size_len_));
            have_space = true;
        } else if (m.type == 0x0003) { // datatype
            if (m.flags & 0x02) return fail("dataset \"" + name + "\" uses a committed (shared) datatype, which is not supported");
            std::string why;
            if (!parse_datatype(d, n, di.dtype, di.big_endian, why)) return fail("dataset \"" + name + "\": unsupported " + why);
            have_type = true;
        } else if (m.type == 0x0008 && n >= 2) { // data layout
            const int ver = d[0];
            if (ver == 3 || ver == 4) {
                di.layout = d[1];
                if (di.layout == 0) {
                    const size_t sz = rd16(d + 2);
                    if (4 + sz > n) return fail("bad compact layout");
                    di.compact.assign(d + 4, d + 4 + sz);
                } else if (di.layout == 1) {
                    di.address = rdn(d + 2, size_off_);
                    di.size = rdn(d + 2 + size_off_, size_len_);
                } else if (di.layout == 2 && ver == 3) {
                    const int rank = d[2];
                    di.address = rdn(d + 3, size_off_);
                    for (int i = 0; i < rank; i++) di.chunk.push_back(rd32(d + 3 + size_off_ + 4 * i));
                } else {
                    return fail("dataset \"" + name + "\": layout class " + std::to_string(di.layout) + " of layout version " + std::to_string(ver) + " is not supported");
                }
            } else if (ver == 1 || ver == 2) {
                const int rank = d[1];
                di.layout = d[2];
                size_t q = 8;
                if (di.layout != 0) { di.address = rdn(d + q, size_off_); q += size_off_; }
                if (q + 4 * (size_t)rank > n) return fail("bad layout message");
                std::vector<uint32_t> dims;
                for (int i = 0; i < rank; i++) dims.push_back(rd32(d + q + 4 * i));
                q += 4 * (size_t)rank;
                if (di.layout == 2) di.chunk = dims;
                if (di.layout == 0) {
                    const size_t sz = rd32(d + q);
                    if (q + 4 + sz > n) return fail("bad compact layout");
                    di.compact.assign(d + q + 4, d + q + 4 + sz);
                }
            } else {
                return fail("data layout version " + std::to_string(ver) + " is not supported");
            }
            have_layout = true;
        } else if (m.type == 0x000B && n >= 2) { // filter pipeline
            const int ver = d[0], nf = d[1];
            size_t q = ver == 1 ? 8 : 2;
            for (int i = 0; i < nf; i++) {
                if (q + 8 > n + 2) return fail("bad filter pipeline");
                const int id = rd16(d + q);
                q += 2;
                size_t name_len = 0;
                if (ver == 1 || id >= 256) { name_len = rd16(d + q); q += 2; }
                q += 2; // flags
                const int ncd = rd16(d + q);
                q += 2;
                if (ver == 1) name_len = (name_len + 7) / 8 * 8;
                q += name_len;
                std::vector<uint32_t> cd;
                for (int k = 0; k < ncd && q + 4 <= n; k++, q += 4) cd.push_back(rd32(d + q));
                if (ver == 1 && (ncd & 1)) q += 4;
                di.filters.push_back({id, cd});
            }
        }
    }
    if (!have_space || !have_type || !have_layout) return fail("\"" + name + "\" is not a (simple, numeric) dataset");
    if (di.layout == 1 && di.address != kUndef) di.address += base_;
    if (di.layout == 2 && di.address != kUndef) di.address += base_;
    return true;
}

bool Reader::walk_chunk_btree(uint64_t node, const DatasetInfo &di, std::vector<uint8_t> &raw, int depth)
{
    if (depth > 32) return fail("chunk B-tree too deep");
    uint8_t h[8 + 16];
    if (!pread(node, h, 8 + 2 * size_off_)) return false;
    if (memcmp(h, "TREE", 4) != 0 || h[4] != 1) return fail("bad chunk B-tree node");
    const int level = h[5], n = rd16(h + 6);
    const size_t nd = di.chunk.size(); // rank + 1
    const size_t key = 8 + 8 * nd, ent = key + size_off_;
    std::vector<uint8_t> b(ent * n + key);
    if (!pread(node + 8 + 2 * size_off_, b.data(), b.size())) return false;
    const size_t rank = nd - 1, es = dtype_size(di.dtype);
    uint64_t chunk_elems = 1;
    for (size_t i = 0; i < rank; i++) chunk_elems *= di.chunk[i];
    for (int i = 0; i < n; i++) {
        const uint8_t *k = &b[ent * i];
        const uint64_t child = base_ + rdn(k + key, size_off_);
        if (level > 0) {
            if (!walk_chunk_btree(child, di, raw, depth + 1)) return false;
            continue;
        }
        const uint32_t nbytes = rd32(k), fmask = rd32(k + 4);
        std::vector<uint64_t> off(rank);
        for (size_t d = 0; d < rank; d++) off[d] = rdn(k + 8 + 8 * d, 8);
        std::vector<uint8_t> buf(nbytes);
        if (!pread(child, buf.data(), nbytes)) return false;
        for (int fi = (int)di.filters.size() - 1; fi >= 0; fi--) { // undo the pipeline back to front
            if (fmask & (1u << fi)) continue;
            const int id = di.filters[fi].first;
            if (id == 3) { // fletcher32: checksum trails the data
                if (buf.size() >= 4) buf.resize(buf.size() - 4);
            } else if (id == 1) { // deflate
                std::vector<uint8_t> o((size_t)chunk_elems * es);
                uLongf olen = (uLongf)o.size();
                if (uncompress(o.data(), &olen, buf.data(), (uLong)buf.size()) != Z_OK) return fail("deflate: corrupt chunk");
                o.resize(olen);
                buf.swap(o);
            } else if (id == 2) { // shuffle: byte planes back to elements
                const size_t e = di.filters[fi].second.empty() ? es : di.filters[fi].second[0];
                if (e > 1 && buf.size() % e == 0) {
                    const size_t ne = buf.size() / e;
                    std::vector<uint8_t> o(buf.size());
                    for (size_t j = 0; j < e; j++)
                        for (size_t x = 0; x < ne; x++) o[x * e + j] = buf[j * ne + x];
                    buf.swap(o);
                }
            } else {
                return fail("filter id " + std::to_string(id) + " is not supported (deflate, shuffle, fletcher32 are)");
            }
        }
        if (buf.size() < chunk_elems * es) return fail("chunk is smaller than its declared shape");
        // copy the chunk into the array, clipping at the dataset's edges; rows along the last dimension are contiguous
        if (rank == 0) { memcpy(raw.data(), buf.data(), es); continue; }
        const uint64_t last_c = di.chunk[rank - 1];
        if (off[rank - 1] >= di.dims[rank - 1]) continue;
        const uint64_t row = std::min<uint64_t>(last_c, di.dims[rank - 1] - off[rank - 1]);
        const uint64_t n_rows = chunk_elems / last_c;
        std::vector<uint64_t> idx(rank, 0);
        for (uint64_t r = 0; r < n_rows; r++) {
            bool inside = true;
            uint64_t dst = 0;
            for (size_t d = 0; d + 1 < rank; d++) {
                const uint64_t g = off[d] + idx[d];
                if (g >= di.dims[d]) { inside = false; break; }
                dst = dst * di.dims[d] + g;
            }
            if (inside) {
                dst = dst * di.dims[rank - 1] + off[rank - 1];
                memcpy(&raw[dst * es], &buf[r * last_c * es], row * es);
            }
            for (int d = (int)rank - 2; d >= 0; d--) { // next row of the chunk
                if (++idx[d] < di.chunk[d]) break;
                idx[d] = 0;
            }
        }
    }
    return true;
}

bool Reader::read_chunked(const DatasetInfo &di, std::vector<uint8_t> &raw)
{
    if (di.chunk.size() != di.dims.size() + 1) return fail("chunk rank does not match the dataspace");
    if (di.address == kUndef) return true; // nothing was ever written: fill value (0)
    return walk_chunk_btree(di.address, di, raw, 0);
}

bool Reader::read(const std::string &name, DType as, void *dst, uint64_t dst_elems)
{
    DatasetInfo di;
    if (!info(name, di)) return false;
    const uint64_t n = di.count();
    if (n != dst_elems) return fail("dataset \"" + name + "\" has different size " + std::to_string(n) + " vs " + std::to_string(dst_elems));
    const size_t es = dtype_size(di.dtype);
    if (n == 0) return true;
    if (di.layout == 1) {
        if (di.address == kUndef) { memset(dst, 0, n * dtype_size(as)); return true; } // never written: fill value
        if (di.dtype == as && !di.big_endian) return pread(di.address, dst, n * es);   // no conversion: straight into the caller's buffer
        const uint64_t step = 1u << 22; // convert in slabs
        std::vector<uint8_t> buf((size_t)std::min(n, step) * es);
        for (uint64_t o = 0; o < n; o += step) {
            const uint64_t m = std::min(step, n - o);
            if (!pread(di.address + o * es, buf.data(), m * es)) return false;
            convert(buf.data(), di.dtype, di.big_endian, as, static_cast<uint8_t *>(dst) + o * dtype_size(as), m);
        }
        return true;
    }
    std::vector<uint8_t> raw;
    if (di.layout == 0) {
        if (di.compact.size() < n * es) return fail("compact dataset is smaller than its dataspace");
        raw = di.compact;
    } else if (di.layout == 2) {
        raw.assign(n * es, 0);
        if (!read_chunked(di, raw)) return false;
    } else {
        return fail("unsupported layout class");
    }
    convert(raw.data(), di.dtype, di.big_endian, as, dst, n);
    return true;
}

// =====================================================================================================================
// Writer
// =====================================================================================================================
void Writer::add(const std::string &name, const std::vector<uint64_t> &dims, DType t, const void *data)
{
    items_.push_back({name, dims, t, data});
}

bool Writer::close()
{
    const int kLeafK = 4, kInternalK = 16; // libhdf5 defaults
    std::vector<size_t> order(items_.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return items_[a].name < items_[b].name; }); // SNOD entries are sorted by name
    for (size_t i = 1; i < order.size(); i++)
        if (items_[order[i]].name == items_[order[i - 1]].name) { err_ = "duplicate dataset name " + items_[order[i]].name; return false; }
    const size_t per_node = 2 * kLeafK;
    const size_t n_snod = std::max<size_t>(1, (items_.size() + per_node - 1) / per_node);
    if (n_snod > 2 * (size_t)kInternalK) { err_ = "too many datasets for a single-level group B-tree"; return false; }

    // ---- local heap data segment: "" at 0, then the names (NUL-terminated, 8-byte aligned) ----
    std::vector<uint8_t> heap(8, 0);
    std::vector<uint64_t> name_off(items_.size());
    for (size_t i : order) {
        name_off[i] = heap.size();
        heap.insert(heap.end(), items_[i].name.begin(), items_[i].name.end());
        heap.push_back(0);
        pad8(heap);
    }
    // a free block at the end (next = 1 means "last", then its size), like the library leaves one
    const uint64_t free_off = heap.size();
    put(heap, 1, 8);
    put(heap, 32, 8);
    heap.resize(heap.size() + 16, 0);

    // ---- dataset object headers ----
    struct Obj { std::vector<uint8_t> bytes; size_t layout_addr_pos; uint64_t data_bytes; };
    std::vector<Obj> objs(items_.size());
    for (size_t i = 0; i < items_.size(); i++) {
        const Item &it = items_[i];
        std::vector<uint8_t> m; // message area
        auto msg = [&](uint16_t type, uint8_t flags, const std::vector<uint8_t> &body) {
            std::vector<uint8_t> b = body;
            pad8(b);
            put(m, type, 2);
            put(m, b.size(), 2);
            m.push_back(flags);
            m.insert(m.end(), 3, 0);
            m.insert(m.end(), b.begin(), b.end());
        };
        std::vector<uint8_t> b;
        // dataspace, version 1
        b = {1, (uint8_t)it.dims.size(), 0, 0, 0, 0, 0, 0};
        for (uint64_t d : it.dims) put(b, d, 8);
        msg(0x0001, 0, b);
        // datatype, version 1
        const size_t es = dtype_size(it.t);
        b.clear();
        if (it.t == DType::F32 || it.t == DType::F64) {
            b = {0x11, 0x20, (uint8_t)(8 * es - 1), 0};
            put(b, es, 4);
            put(b, 0, 2);
            put(b, 8 * es, 2);
            if (es == 4) { b.insert(b.end(), {23, 8, 0, 23}); put(b, 127, 4); }
            else { b.insert(b.end(), {52, 11, 0, 52}); put(b, 1023, 4); }
        } else {
            const bool sgn = it.t == DType::I8 || it.t == DType::I16 || it.t == DType::I32 || it.t == DType::I64;
            b = {0x10, (uint8_t)(sgn ? 0x08 : 0x00), 0, 0};
            put(b, es, 4);
            put(b, 0, 2);
            put(b, 8 * es, 2);
        }
        msg(0x0003, 1, b);
        // fill value, version 2: allocate late, write fill if set, default fill value
        b = {2, 2, 2, 1, 0, 0, 0, 0};
        msg(0x0005, 1, b);
        // data layout, version 3, contiguous
        uint64_t nbytes = es;
        for (uint64_t d : it.dims) nbytes *= d;
        b = {3, 1};
        const size_t addr_in_body = b.size();
        put(b, kUndef, 8);
        put(b, nbytes, 8);
        const size_t layout_pos = m.size() + 8 + addr_in_body;
        msg(0x0008, 0, b);

        Obj &o = objs[i];
        o.bytes = {1, 0};
        put(o.bytes, 4, 2);        // number of messages
        put(o.bytes, 1, 4);        // reference count
        put(o.bytes, m.size(), 4); // header data size
        put(o.bytes, 0, 4);        // alignment
        o.layout_addr_pos = o.bytes.size() + layout_pos;
        o.bytes.insert(o.bytes.end(), m.begin(), m.end());
        o.data_bytes = nbytes;
    }

    // ---- addresses ----
    uint64_t at = 96; // superblock v0 with 8-byte offsets: 56 + 40
    const uint64_t root_oh = at;
    at += 16 + 24;
    const uint64_t btree = at;
    at += 24 + (2 * kInternalK + 1) * 8 + 2 * kInternalK * 8;
    const uint64_t heap_hdr = at;
    at += 32;
    const uint64_t heap_data = at;
    at += heap.size();
    std::vector<uint64_t> snod(n_snod);
    for (auto &s : snod) { s = at; at += 8 + per_node * 40; }
    std::vector<uint64_t> oh(items_.size()), data(items_.size());
    for (size_t i = 0; i < items_.size(); i++) { oh[i] = at; at += objs[i].bytes.size(); }
    for (size_t i = 0; i < items_.size(); i++) {
        at = (at + 7) / 8 * 8;
        data[i] = at;
        at += objs[i].data_bytes;
    }
    const uint64_t eof = at;

    // ---- metadata block ----
    std::vector<uint8_t> md;
    md.insert(md.end(), kSig, kSig + 8);
    md.insert(md.end(), {0, 0, 0, 0, 0, 8, 8, 0}); // superblock v0, free-space v0, root STE v0, -, shared header v0, offsets 8, lengths 8, -
    put(md, kLeafK, 2);
    put(md, kInternalK, 2);
    put(md, 0, 4);      // file consistency flags
    put(md, 0, 8);      // base address
    put(md, kUndef, 8); // free-space info
    put(md, eof, 8);    // end of file
    put(md, kUndef, 8); // driver info
    put(md, 0, 8);      // root entry: link name offset
    put(md, root_oh, 8);
    put(md, 1, 4);      // cache type 1: scratch pad holds B-tree and heap addresses
    put(md, 0, 4);
    put(md, btree, 8);
    put(md, heap_hdr, 8);
    // root group object header: one symbol-table message
    md.insert(md.end(), {1, 0});
    put(md, 1, 2);
    put(md, 1, 4);
    put(md, 24, 4);
    put(md, 0, 4);
    put(md, 0x0011, 2);
    put(md, 16, 2);
    md.insert(md.end(), {0, 0, 0, 0});
    put(md, btree, 8);
    put(md, heap_hdr, 8);
    // group B-tree node (type 0, leaf level): key[0] = "" ; key[i+1] = largest name of child i
    md.insert(md.end(), {'T', 'R', 'E', 'E', 0, 0});
    put(md, items_.empty() ? 0 : n_snod, 2);
    put(md, kUndef, 8);
    put(md, kUndef, 8);
    {
        std::vector<uint8_t> kc;
        put(kc, 0, 8);
        for (size_t s = 0; s < n_snod && !items_.empty(); s++) {
            const size_t last = std::min(items_.size(), (s + 1) * per_node) - 1;
            put(kc, snod[s], 8);
            put(kc, name_off[order[last]], 8);
        }
        kc.resize((2 * kInternalK + 1) * 8 + 2 * kInternalK * 8, 0);
        md.insert(md.end(), kc.begin(), kc.end());
    }
    // local heap header + data segment
    md.insert(md.end(), {'H', 'E', 'A', 'P', 0, 0, 0, 0});
    put(md, heap.size(), 8);
    put(md, free_off, 8);
    put(md, heap_data, 8);
    md.insert(md.end(), heap.begin(), heap.end());
    // symbol table nodes
    for (size_t s = 0; s < n_snod; s++) {
        const size_t first = s * per_node, last = std::min(items_.size(), first + per_node);
        md.insert(md.end(), {'S', 'N', 'O', 'D', 1, 0});
        put(md, last > first ? last - first : 0, 2);
        for (size_t k = first; k < first + per_node; k++) {
            if (k < last) {
                put(md, name_off[order[k]], 8);
                put(md, oh[order[k]], 8);
            } else {
                put(md, 0, 16);
            }
            put(md, 0, 8);  // cache type 0 + reserved
            put(md, 0, 8);  // scratch pad
            put(md, 0, 8);
        }
    }
    // dataset object headers, layout addresses patched in
    for (size_t i = 0; i < items_.size(); i++) {
        std::vector<uint8_t> b = objs[i].bytes;
        const uint64_t a = objs[i].data_bytes ? data[i] : kUndef;
        for (int k = 0; k < 8; k++) b[objs[i].layout_addr_pos + k] = (uint8_t)(a >> (8 * k));
        md.insert(md.end(), b.begin(), b.end());
    }

    FILE *f = fopen(path_.c_str(), "wb");
    if (!f) { err_ = "cannot create " + path_; return false; }
    bool ok = fwrite(md.data(), 1, md.size(), f) == md.size();
    uint64_t pos = md.size();
    static const uint8_t zeros[8] = {0};
    for (size_t i = 0; i < items_.size() && ok; i++) {
        if (data[i] > pos) { ok = fwrite(zeros, 1, (size_t)(data[i] - pos), f) == data[i] - pos; pos = data[i]; }
        if (ok && objs[i].data_bytes) {
            ok = fwrite(items_[i].data, 1, (size_t)objs[i].data_bytes, f) == objs[i].data_bytes;
            pos += objs[i].data_bytes;
        }
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) err_ = "write error on " + path_;
    return ok && pos == eof;
}

} // namespace h5
} // namespace swk_host
