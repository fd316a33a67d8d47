// raptor_b200/csrc/h5_io.cu -- minimal HDF5 reader for rl-tools' `checkpoint.h5`, host code only (no libhdf5 / HighFive in this stack).
//
// rl::loop::steps::checkpoint::save (rl/loop/steps/checkpoint/operations_cpu.h:119-160) writes, through HighFive with the library's default
// ("earliest") file format:
//     /actor                          group; attributes checkpoint_name, meta (environment JSON), type = "sequential"
//     /actor/layers/<k>               one group per layer (nn_models/sequential/persist.h:14-21); attributes type, activation_function
//     /actor/layers/<k>/<param>/parameters     dataset per parameter (nn/parameters/persist.h:10-13; containers/{matrix,tensor}/persist.h)
//     /example/input, /example/output          the known-answer pair
// That format is: superblock version 0 (or 1), version-1 object headers with continuation blocks, "old style" groups (symbol-table message ->
// version-1 B-tree of symbol-table nodes + local heap for the names), contiguous (or compact) dataset layout, IEEE float datatypes, and
// variable-length string attributes stored in global heap collections.  This unit reads exactly that subset from a memory image and
// reports everything else (version-2 object headers / link messages of `libver=latest` files, chunked or filtered datasets, ...) as an error
// naming what it met.  Every access is bounds-checked; a truncated or corrupted file gives an error, never an out-of-range read.
// Structure layouts follow the published HDF5 File Format Specification (version 1.1 / 2.0: sections II.A disk format level 0A superblock,
// III.A B-trees, III.B symbol-table nodes, III.D local heaps, III.E global heaps, IV.A object headers and messages 0x1 / 0x3 / 0x8 / 0xC /
// 0x10 / 0x11).
#include "h5_io.h"

#include <cstring>
#include <set>

namespace b200l2f {
namespace {

struct Reader {
    const unsigned char* d; size_t n;
    H5Contents& out; std::string& err;
    int so = 8, sl = 8;                 // size of offsets / lengths
    uint64_t base = 0;                  // base address: every file address is relative to it
    std::set<uint64_t> visited;         // object headers already walked (hard links may form cycles)
    uint64_t tree_nodes = 0;            // B-tree nodes walked (bounds the work a corrupted, self-referencing tree can cause)

    bool fail(const std::string& m){ if(err.empty()) err = m; return false; }
    bool in(uint64_t off, uint64_t len) const { return off <= n && len <= n - off; }
    // little-endian unsigned of `bytes` bytes; the caller has checked the range
    uint64_t le(uint64_t off, int bytes) const { uint64_t v = 0; for(int i = bytes - 1; i >= 0; i--) v = (v << 8) | d[off + (uint64_t)i]; return v; }
    bool rd(uint64_t off, int bytes, uint64_t& v){ if(!in(off, (uint64_t)bytes)) return fail("read past the end of the file (offset " + std::to_string(off) + ")"); v = le(off, bytes); return true; }
    bool undefined(uint64_t a, int bytes) const { return bytes >= 8 ? a == ~0ull : a == ((1ull << (8 * bytes)) - 1); }
    bool addr(uint64_t off, uint64_t& a){ if(!rd(off, so, a)) return false; if(!undefined(a, so)) a += base; return true; }
    bool sig(uint64_t off, const char* s4){ return in(off, 4) && std::memcmp(d + off, s4, 4) == 0; }

    // ---- superblock ---------------------------------------------------------------------------------------------------------------------
    bool superblock(uint64_t& root_header){
        static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t at = 0; bool found = false;
        for(uint64_t o = 0; o + 8 <= n; o = o ? o * 2 : 512){ if(std::memcmp(d + o, magic, 8) == 0){ at = o; found = true; break; } }
        if(!found) return fail("not an HDF5 file (no signature)");
        uint64_t v;
        if(!rd(at + 8, 1, v)) return false;
        if(v > 1) return fail("superblock version " + std::to_string(v) + " (a `libver=latest` file): only the default file format rl-tools / HighFive write is read");
        uint64_t a, b;
        if(!rd(at + 13, 1, a) || !rd(at + 14, 1, b)) return false;
        so = (int)a; sl = (int)b;
        if(so != 8 || sl != 8) return fail("size of offsets / lengths " + std::to_string(so) + " / " + std::to_string(sl) + ": only the 8-byte sizes libhdf5 writes are read");
        uint64_t p = at + 24 + (v == 1 ? 4 : 0);                      // version 1 adds indexed-storage K + 2 reserved bytes
        if(!rd(p, so, base)) return false;
        if(undefined(base, so)) base = 0;
        p += 4ull * so;                                               // base, free-space info, end of file, driver info
        // root group symbol-table entry: link name offset, object header address, cache type, reserved, scratch
        return addr(p + so, root_header);
    }

    // ---- object header (version 1) ------------------------------------------------------------------------------------------------------
    struct Msg { unsigned type; uint64_t at; uint64_t size; unsigned flags; };
    bool messages(uint64_t header, std::vector<Msg>& msgs){
        if(sig(header, "OHDR")) return fail("version-2 object header (a `libver=latest` file) is not supported");
        uint64_t ver, count, size;
        if(!rd(header, 1, ver) || !rd(header + 2, 2, count) || !rd(header + 8, 4, size)) return false;
        if(ver != 1) return fail("object header version " + std::to_string(ver) + " at " + std::to_string(header));
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{header + 16, size}};
        for(size_t bi = 0; bi < blocks.size() && msgs.size() < count; bi++){
            uint64_t p = blocks[bi].first; const uint64_t len = blocks[bi].second;
            if(!in(p, len)) return fail("object header block outside the file");
            const uint64_t end = p + len;
            while(p + 8 <= end && msgs.size() < count){
                Msg m{(unsigned)le(p, 2), p + 8, le(p + 2, 2), (unsigned)le(p + 4, 1)};
                if(m.at + m.size > end) return fail("object header message overruns its block");
                if(m.type == 0x10){                                   // continuation: offset, length
                    if(m.size < (uint64_t)(so + sl)) return fail("short continuation message");
                    uint64_t a, l; if(!addr(m.at, a) || !rd(m.at + so, sl, l)) return false;
                    if(blocks.size() > 4096) return fail("too many object header continuation blocks");
                    blocks.push_back({a, l});
                }
                msgs.push_back(m);
                p = m.at + m.size;
            }
        }
        return true;
    }

    // ---- groups: version-1 B-tree of symbol-table nodes, names in a local heap -------------------------------------------------------------------
    bool heap_name(uint64_t heap, uint64_t off, std::string& name){
        if(!sig(heap, "HEAP")) return fail("local heap signature missing");
        uint64_t seg_size, seg;
        if(!rd(heap + 8, sl, seg_size) || !addr(heap + 8 + 2ull * sl, seg)) return false;
        if(!in(seg, seg_size) || off >= seg_size) return fail("link name outside the local heap");
        const unsigned char* s = d + seg + off; const size_t room = (size_t)(seg_size - off);
        const void* z = std::memchr(s, 0, room);
        if(!z) return fail("unterminated link name");
        name.assign((const char*)s, (const char*)z);
        return true;
    }
    bool btree(uint64_t node, uint64_t heap, int depth, std::vector<std::pair<std::string, uint64_t>>& links){
        if(depth > 16 || ++tree_nodes > (1u << 20)) return fail("group B-tree too deep / too large");
        if(!sig(node, "TREE")) return fail("B-tree node signature missing");
        uint64_t type, level, used;
        if(!rd(node + 4, 1, type) || !rd(node + 5, 1, level) || !rd(node + 6, 2, used)) return false;
        if(type != 0) return fail("B-tree node is not a group node");
        uint64_t p = node + 8 + 2ull * so;                            // past the sibling addresses; then key 0, child 0, key 1, ...
        for(uint64_t i = 0; i < used; i++){
            uint64_t child; if(!addr(p + sl, child)) return false;
            p += (uint64_t)(sl + so);
            if(level > 0){ if(!btree(child, heap, depth + 1, links)) return false; continue; }
            if(!sig(child, "SNOD")) return fail("symbol-table node signature missing");
            uint64_t count; if(!rd(child + 6, 2, count)) return false;
            uint64_t q = child + 8;
            for(uint64_t k = 0; k < count; k++, q += 2ull * so + 24){
                uint64_t name_off, header; if(!rd(q, so, name_off) || !addr(q + so, header)) return false;
                std::string name; if(!heap_name(heap, name_off, name)) return false;
                if(links.size() > (1u << 20)) return fail("too many links");
                links.push_back({name, header});
            }
        }
        return true;
    }

    // ---- datatype / dataspace ---------------------------------------------------------------------------------------------------------------
    struct Type { int cls = -1; uint64_t size = 0; bool big_endian = false, is_signed = false, vlen_string = false; };
    bool datatype(uint64_t at, uint64_t room, Type& t){
        if(room < 8 || !in(at, 8)) return fail("short datatype message");
        t.cls = (int)(d[at] & 15); t.size = le(at + 4, 4);
        const unsigned bits0 = d[at + 1];
        if(t.cls == 1){                                               // floating point: properties = bit offset, precision, exponent location / size, mantissa location / size, bias
            t.big_endian = bits0 & 1;
            if(room < 20 || !in(at, 20)) return fail("short floating-point datatype");
            const unsigned exp_size = d[at + 13], man_size = d[at + 15];
            if(!((t.size == 4 && exp_size == 8 && man_size == 23) || (t.size == 8 && exp_size == 11 && man_size == 52))) return fail("floating-point datatype is not IEEE binary32 / binary64");
        }
        else if(t.cls == 0){ t.big_endian = bits0 & 1; t.is_signed = bits0 & 8; if(t.size != 1 && t.size != 2 && t.size != 4 && t.size != 8) return fail("integer datatype of unsupported size"); }
        else if(t.cls == 9){ t.vlen_string = (bits0 & 15) == 1; }
        return true;
    }
    bool dataspace(uint64_t at, uint64_t room, std::vector<int64_t>& dims){
        if(room < 4 || !in(at, 4)) return fail("short dataspace message");
        const unsigned ver = d[at], rank = d[at + 1];
        if(ver != 1 && ver != 2) return fail("dataspace version " + std::to_string(ver));
        if(rank > 32) return fail("dataspace rank " + std::to_string(rank));
        const uint64_t p = at + (ver == 1 ? 8 : 4);
        if(room < (p - at) + (uint64_t)rank * sl || !in(p, (uint64_t)rank * sl)) return fail("short dataspace message");
        dims.clear();
        for(unsigned i = 0; i < rank; i++) dims.push_back((int64_t)le(p + (uint64_t)i * sl, sl));
        return true;
    }
    uint64_t be_or_le(uint64_t off, int bytes, bool big) const {
        if(!big) return le(off, bytes);
        uint64_t v = 0; for(int i = 0; i < bytes; i++) v = (v << 8) | d[off + (uint64_t)i]; return v;
    }
    bool convert(const Type& t, uint64_t at, uint64_t count, std::vector<float>& data){
        if(count > (1ull << 32) || !in(at, count * t.size)) return fail("dataset data outside the file");
        data.resize((size_t)count);
        for(uint64_t i = 0; i < count; i++){
            const uint64_t raw = be_or_le(at + i * t.size, (int)t.size, t.big_endian);
            float f;
            if(t.cls == 1){
                if(t.size == 4){ const uint32_t u = (uint32_t)raw; std::memcpy(&f, &u, 4); }
                else { double x; std::memcpy(&x, &raw, 8); f = (float)x; }
            }
            else if(t.is_signed){ const int sh = 64 - 8 * (int)t.size; f = (float)((int64_t)(raw << sh) >> sh); }
            else f = (float)raw;
            data[(size_t)i] = f;
        }
        return true;
    }

    // ---- global heap (variable-length attribute strings) ---------------------------------------------------------------------------------------
    bool global_heap_object(uint64_t collection, uint64_t index, uint64_t& at, uint64_t& size){
        if(!sig(collection, "GCOL")) return fail("global heap collection signature missing");
        uint64_t total; if(!rd(collection + 8, sl, total)) return false;
        if(!in(collection, total)) return fail("global heap collection outside the file");
        const uint64_t end = collection + total;                      // in() checked: no wrap
        const uint64_t hdr = 8 + (uint64_t)sl;                        // object header: index (2) | reference count (2) | reserved (4) | size
        uint64_t p = collection + hdr;
        while(end - p >= hdr){
            const uint64_t idx = le(p, 2), sz = le(p + 8, sl);
            // the size comes from the file: compare by subtraction so that a value near 2^64 can neither wrap the bound nor the stride
            if(sz > end - (p + hdr)) return fail("global heap object overruns its collection");
            if(idx == index){ at = p + hdr; size = sz; return true; }
            if(idx == 0) break;                                       // object 0 = the free space at the end
            const uint64_t step = hdr + ((sz + 7) & ~7ull);           // sz <= end - p - hdr < 2^63: the rounding cannot wrap
            if(step > end - p) break;
            p += step;
        }
        return fail("global heap object " + std::to_string(index) + " not found");
    }

    // ---- attribute message ---------------------------------------------------------------------------------------------------------------------
    bool attribute(const Msg& m, const std::string& object){
        if(m.size < 8) return fail("short attribute message");
        const unsigned ver = d[m.at];
        if(ver < 1 || ver > 3) return fail("attribute message version " + std::to_string(ver));
        if(ver >= 2 && (d[m.at + 1] & 3)) return true;                // shared datatype / dataspace: not something rl-tools writes; skipped
        const uint64_t name_size = le(m.at + 2, 2), type_size = le(m.at + 4, 2), space_size = le(m.at + 6, 2);
        auto pad = [&](uint64_t x){ return ver == 1 ? (x + 7) & ~7ull : x; };
        uint64_t p = m.at + 8 + (ver == 3 ? 1 : 0);
        const uint64_t end = m.at + m.size;
        if(p + pad(name_size) + pad(type_size) + pad(space_size) > end) return fail("attribute message overruns");
        std::string name((const char*)d + p, (size_t)name_size); name = name.substr(0, name.find('\0')); p += pad(name_size);
        Type t; if(!datatype(p, type_size, t)) return false; p += pad(type_size);
        std::vector<int64_t> dims; if(!dataspace(p, space_size, dims)) return false; p += pad(space_size);
        uint64_t count = 1; for(auto v : dims) count *= (uint64_t)v;
        if(count != 1) return true;                                   // arrays of strings: skipped
        std::string value;
        if(t.cls == 3){                                               // fixed-length string
            if(p + t.size > end) return fail("attribute data overruns");
            value.assign((const char*)d + p, (size_t)t.size); value = value.substr(0, value.find('\0'));
        }
        else if(t.cls == 9 && t.vlen_string){                         // length, global heap collection address, object index
            if(p + 4 + (uint64_t)so + 4 > end) return fail("attribute data overruns");
            uint64_t collection, at, size; const uint64_t index = le(p + 4 + (uint64_t)so, 4);
            if(!addr(p + 4, collection)) return false;
            if(undefined(collection, so) || collection == base){ value.clear(); }   // null / empty string
            else { if(!global_heap_object(collection, index, at, size)) return false; value.assign((const char*)d + at, (size_t)size); value = value.substr(0, value.find('\0')); }
        }
        else return true;                                             // numeric attribute: nothing in a checkpoint needs it
        out.attributes.push_back({object, name, value});
        return true;
    }

    // ---- one object: group (recurse) or dataset ----------------------------------------------------------------------------------------------------
    bool object(uint64_t header, const std::string& path, int depth){
        if(depth > 64) return fail("group nesting too deep");
        if(!visited.insert(header).second) return true;
        std::vector<Msg> msgs; if(!messages(header, msgs)) return false;
        const std::string self = path.empty() ? "/" : path;
        const Msg *group = nullptr, *space = nullptr, *type = nullptr, *layout = nullptr;
        for(auto& m : msgs){
            if(m.type == 0x11) group = &m; else if(m.type == 0x1) space = &m; else if(m.type == 0x3) type = &m; else if(m.type == 0x8) layout = &m;
            else if(m.type == 0x2 || m.type == 0x6) return fail("new-style group (link messages) at '" + self + "' is not supported");
            else if(m.type == 0xB) return fail("filtered dataset '" + self + "' is not supported");
        }
        for(auto& m : msgs) if(m.type == 0xC && !attribute(m, self)) return false;
        if(group){
            if(group->size < 2ull * so) return fail("short symbol-table message");
            uint64_t tree, heap; if(!addr(group->at, tree) || !addr(group->at + so, heap)) return false;
            out.groups.push_back(self);
            std::vector<std::pair<std::string, uint64_t>> links;
            if(!btree(tree, heap, 0, links)) return false;
            for(auto& l : links) if(!object(l.second, path + "/" + l.first, depth + 1)) return false;
            return true;
        }
        if(!(space && type && layout)) return true;                   // committed datatype or something else without data
        if(type->flags & 2) return true;                              // shared (committed) datatype: skipped
        Type t; if(!datatype(type->at, type->size, t)) return false;
        if(t.cls != 0 && t.cls != 1) return true;                     // only numeric datasets become tensors
        H5Dataset ds; ds.path = self; ds.elem_size = (int)t.size;
        if(!dataspace(space->at, space->size, ds.dims)) return false;
        uint64_t count = 1; for(auto v : ds.dims){ if(v < 0 || (v && count > (1ull << 32) / (uint64_t)v)) return fail("dataset '" + self + "' is too large"); count *= (uint64_t)v; }
        if(layout->size < 2) return fail("short layout message");
        const unsigned lver = d[layout->at];
        uint64_t data_at = 0, data_size = 0; unsigned cls;
        if(lver == 3){
            cls = d[layout->at + 1];
            if(cls == 0){ if(layout->size < 4) return fail("short layout message"); data_size = le(layout->at + 2, 2); data_at = layout->at + 4; if(data_at + data_size > layout->at + layout->size) return fail("compact data overruns its message"); }
            else if(cls == 1){ if(layout->size < 2 + (uint64_t)so + sl) return fail("short layout message"); if(!addr(layout->at + 2, data_at) || !rd(layout->at + 2 + so, sl, data_size)) return false; }
        }
        else if(lver == 1 || lver == 2){
            if(layout->size < 8) return fail("short layout message");
            const unsigned rank = d[layout->at + 1]; cls = d[layout->at + 2];
            uint64_t p = layout->at + 8;
            if(cls == 1){ if(layout->size < 8 + (uint64_t)so) return fail("short layout message"); if(!addr(p, data_at)) return false; data_size = count * t.size; }
            else if(cls == 0){ p += 4ull * rank; if(p + 4 > layout->at + layout->size) return fail("short layout message"); data_size = le(p, 4); data_at = p + 4; if(data_at + data_size > layout->at + layout->size) return fail("compact data overruns its message"); }
        }
        else return fail("layout message version " + std::to_string(lver) + " of '" + self + "'");
        if(cls == 2) return fail("chunked dataset '" + self + "' is not supported (rl-tools writes contiguous datasets)");
        if(cls > 2) return fail("dataset '" + self + "' has layout class " + std::to_string(cls));
        if(count){
            if(cls == 1 && undefined(data_at, so)) return fail("dataset '" + self + "' has no data allocated");
            if(data_size < count * t.size) return fail("dataset '" + self + "' holds fewer bytes than its shape needs");
            if(!convert(t, data_at, count, ds.data)) return false;
        }
        out.datasets.push_back(std::move(ds));
        return true;
    }
};

}  // namespace

bool h5_has_signature(const unsigned char* bytes, size_t length){
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if(!bytes) return false;
    for(uint64_t o = 0; o + 8 <= length; o = o ? o * 2 : 512) if(std::memcmp(bytes + o, magic, 8) == 0) return true;
    return false;
}

bool h5_read(const unsigned char* bytes, size_t length, H5Contents& out, std::string& err){
    out = H5Contents(); err.clear();
    if(!bytes) { err = "null buffer"; return false; }
    Reader r{bytes, length, out, err};
    uint64_t root;
    if(!r.superblock(root)) return false;
    return r.object(root, "", 0);
}

}  // namespace b200l2f
