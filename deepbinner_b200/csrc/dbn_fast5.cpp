// Native fast5 reader: read_id + raw signal of single-read fast5 files with the
// semantics of reference load_fast5s.py:25-49 (get_read_id_and_signal) and :93-98
// (get_root_level_keys), without libhdf5 (absent from the image).  Implements the HDF5 subset those
// files use: superblock v0-v3, object headers v1 (continuations) and v2,
// old-style groups (B-tree v1 + SNOD + local heap), compact link messages, dense links (fractal heap
// direct blocks addressed through a single-leaf v2 B-tree name index), fixed-length string
// attributes, int16 datasets with contiguous / compact / chunked (B-tree v1) layout and the deflate
// (+ shuffle) filters.  A batch entry point parses many files on a thread pool.  Host-only code.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/deepbinner_b200.h"
#include "dbn_inflate.h"

namespace dbn {
int fail(int code, const char* fmt, ...);
}

namespace {

constexpr uint64_t kUndef = 0xFFFFFFFFFFFFFFFFull;

struct ParseError {
    const char* what;
};

struct Buf {
    const uint8_t* p = nullptr;
    size_t n = 0;
    void need(uint64_t off, uint64_t len) const {
        if (off > n || len > n - off) throw ParseError{"read beyond end of file"};
    }
    uint8_t u8(uint64_t o) const { need(o, 1); return p[o]; }
    uint16_t u16(uint64_t o) const { need(o, 2); uint16_t v; std::memcpy(&v, p + o, 2); return v; }
    uint32_t u32(uint64_t o) const { need(o, 4); uint32_t v; std::memcpy(&v, p + o, 4); return v; }
    uint64_t u64(uint64_t o) const { need(o, 8); uint64_t v; std::memcpy(&v, p + o, 8); return v; }
    uint64_t uvar(uint64_t o, int bytes) const {
        need(o, bytes);
        uint64_t v = 0;
        for (int i = 0; i < bytes; ++i) v |= static_cast<uint64_t>(p[o + i]) << (8 * i);
        return v;
    }
    bool sig(uint64_t o, const char* s) const { return o + 4 <= n && std::memcmp(p + o, s, 4) == 0; }
};

struct Message {
    int type;
    uint64_t off;   // offset of the message body in the file
    uint32_t size;
};

struct Link {
    std::string name;
    uint64_t addr;
};

// zlib inflate of one whole deflate stream into a caller buffer; returns the bytes produced (a stream
// longer than the buffer is cut there: edge chunks are declared larger than the data set).  The z_stream
// is kept per thread and reset between calls (inflateInit allocates and initialises ~40 KB every time).
size_t zlib_inflate_into(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len) {
    struct Stream {
        z_stream zs{};
        bool ok = false;
        Stream() { ok = inflateInit(&zs) == Z_OK; }
        ~Stream() { if (ok) inflateEnd(&zs); }
    };
    thread_local Stream st;
    if (!st.ok || inflateReset(&st.zs) != Z_OK) throw ParseError{"zlib init failed"};
    st.zs.next_in = const_cast<Bytef*>(src);
    st.zs.avail_in = static_cast<uInt>(src_len);
    st.zs.next_out = dst;
    st.zs.avail_out = static_cast<uInt>(dst_len);
    const int rc = inflate(&st.zs, Z_FINISH);
    if (rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) throw ParseError{"inflate failed"};
    return st.zs.total_out;
}

// The signal chunk goes through the word-at-a-time decoder of dbn_inflate.h (2-3x zlib on this
// literal-heavy data); a stream it rejects is handed to zlib, whose verdict stands.  `src` is readable
// for kFilePadding bytes beyond its end (read_file pads the file buffer).
constexpr size_t kFilePadding = 32;
bool g_use_zlib_only = getenv("DEEPBINNER_B200_ZLIB") != nullptr;
size_t inflate_into(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len) {
    size_t got = 0;
    if (!g_use_zlib_only && dbn_inflate::inflate_zlib(src, src_len, dst, dst_len, &got)) return got;
    return zlib_inflate_into(src, src_len, dst, dst_len);
}

class File {
  public:
    File(const std::vector<uint8_t>& data, size_t size) {
        b_.p = data.data();
        b_.n = size;
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        if (b_.n < 96 || std::memcmp(b_.p, sig, 8) != 0) throw ParseError{"not an HDF5 file"};
        const int version = b_.u8(8);
        if (version == 0 || version == 1) {
            if (b_.u8(13) != 8 || b_.u8(14) != 8) throw ParseError{"unsupported offset size"};
            const uint64_t p = version == 0 ? 24 : 28;
            root_ = b_.u64(p + 32 + 8);
        } else if (version == 2 || version == 3) {
            if (b_.u8(9) != 8 || b_.u8(10) != 8) throw ParseError{"unsupported offset size"};
            root_ = b_.u64(36);
        } else {
            throw ParseError{"unsupported superblock version"};
        }
    }

    uint64_t root() const { return root_; }

    std::vector<Message> messages(uint64_t addr) const {
        std::vector<Message> out;
        if (b_.sig(addr, "OHDR")) {
            const int flags = b_.u8(addr + 5);
            uint64_t p = addr + 6;
            if (flags & 0x20) p += 16;
            if (flags & 0x10) p += 4;
            const int sb = 1 << (flags & 3);
            const uint64_t chunk0 = b_.uvar(p, sb);
            p += sb;
            const bool track = flags & 0x04;
            std::vector<std::pair<uint64_t, uint64_t>> blocks{{p, chunk0}};
            for (size_t bi = 0; bi < blocks.size() && bi < 64; ++bi) {
                uint64_t q = blocks[bi].first;
                const uint64_t end = q + blocks[bi].second;
                while (q + 4 <= end) {
                    const int type = b_.u8(q);
                    const uint32_t size = b_.u16(q + 1);
                    q += 4 + (track ? 2 : 0);
                    b_.need(q, size);
                    if (type == 0x10)
                        blocks.push_back({b_.u64(q) + 4, b_.u64(q + 8) - 8});
                    else if (type != 0)
                        out.push_back({type, q, size});
                    q += size;
                }
            }
            return out;
        }
        if (b_.u8(addr) != 1) throw ParseError{"unsupported object header"};
        const int nmsgs = b_.u16(addr + 2);
        const uint32_t hdr = b_.u32(addr + 8);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, hdr}};
        int seen = 0;
        for (size_t bi = 0; bi < blocks.size() && bi < 64 && seen < nmsgs; ++bi) {
            uint64_t q = blocks[bi].first;
            const uint64_t end = q + blocks[bi].second;
            while (q + 8 <= end && seen < nmsgs) {
                const int type = b_.u16(q);
                const uint32_t size = b_.u16(q + 2);
                b_.need(q + 8, size);
                if (type == 0x10) blocks.push_back({b_.u64(q + 8), b_.u64(q + 16)});
                out.push_back({type, q + 8, size});
                q += 8 + size;
                ++seen;
            }
        }
        return out;
    }

    std::vector<Link> links(uint64_t addr) const {
        std::vector<Link> out;
        for (const Message& m : messages(addr)) {
            if (m.type == 0x11) {
                const uint64_t btree = b_.u64(m.off), heap = b_.u64(m.off + 8);
                if (!b_.sig(heap, "HEAP")) throw ParseError{"bad local heap"};
                walk_group(btree, b_.u64(heap + 24), &out, 0);
            } else if (m.type == 0x06) {
                Link l;
                uint64_t end;
                if (parse_link(m.off, m.off + m.size, &l, &end)) out.push_back(l);
            } else if (m.type == 0x02) {
                const int flags = b_.u8(m.off + 1);
                uint64_t p = m.off + 2 + ((flags & 1) ? 8 : 0);
                const uint64_t fheap = b_.u64(p), name_index = b_.u64(p + 8);
                if (fheap != kUndef) dense_links(fheap, name_index, &out);
            }
        }
        return out;
    }

    bool find(uint64_t addr, const std::string& name, uint64_t* child) const {
        for (const Link& l : links(addr))
            if (l.name == name) {
                *child = l.addr;
                return true;
            }
        return false;
    }

    // fixed-length string attribute (read_id)
    bool string_attr(uint64_t addr, const std::string& name, std::string* value) const {
        for (const Message& m : messages(addr)) {
            if (m.type != 0x0C) continue;
            const int version = b_.u8(m.off);
            const uint32_t name_size = b_.u16(m.off + 2), dt_size = b_.u16(m.off + 4),
                           ds_size = b_.u16(m.off + 6);
            uint64_t p = m.off + (version == 3 ? 9 : 8);
            auto pad = [&](uint32_t v) { return version == 1 ? (v + 7) / 8 * 8 : v; };
            b_.need(p, name_size);
            const std::string nm(reinterpret_cast<const char*>(b_.p + p),
                                 strnlen(reinterpret_cast<const char*>(b_.p + p), name_size));
            p += pad(name_size);
            if (nm != name) continue;
            const int cls = b_.u8(p) & 0x0F;
            const uint32_t elem = b_.u32(p + 4);
            if (cls != 3) return false;   // only fixed-length strings are needed here
            p += pad(dt_size) + pad(ds_size);
            b_.need(p, elem);
            value->assign(reinterpret_cast<const char*>(b_.p + p),
                          strnlen(reinterpret_cast<const char*>(b_.p + p), elem));
            return true;
        }
        return false;
    }

    // rank-1 int16 dataset -> out
    // `head` > 0: only the first `head` samples are needed - chunks beyond them are skipped and the inflate of
    // the chunk that holds them stops there (a start-model-only run never looks further into a read).
    // Returns the full length of the data set.
    uint64_t read_i16(uint64_t addr, std::vector<int16_t>* out, uint64_t head = 0) const {
        uint64_t len = 0;
        bool have_space = false, have_layout = false, is_i16 = false;
        Message layout{};
        std::vector<int> filters;
        for (const Message& m : messages(addr)) {
            if (m.type == 0x01) {
                const int version = b_.u8(m.off), rank = b_.u8(m.off + 1);
                if (rank != 1) throw ParseError{"Signal dataset is not rank 1"};
                len = b_.u64(m.off + (version == 1 ? 8 : 4));
                have_space = true;
            } else if (m.type == 0x03) {
                const int cls = b_.u8(m.off) & 0x0F;
                is_i16 = cls == 0 && b_.u32(m.off + 4) == 2 && !(b_.u8(m.off + 1) & 1);
            } else if (m.type == 0x08) {
                layout = m;
                have_layout = true;
            } else if (m.type == 0x0B) {
                const int version = b_.u8(m.off), nf = b_.u8(m.off + 1);
                uint64_t p = m.off + (version == 1 ? 8 : 2);
                for (int i = 0; i < nf; ++i) {
                    const int id = b_.u16(p);
                    p += 2;
                    uint32_t name_len = 0;
                    if (version == 1 || id >= 256) {
                        name_len = b_.u16(p);
                        p += 2;
                    }
                    const uint32_t ncv = b_.u16(p + 2);
                    p += 4;
                    p += version == 1 ? (name_len + 7) / 8 * 8 : name_len;
                    p += 4ull * ncv;
                    if (version == 1 && (ncv & 1)) p += 4;
                    filters.push_back(id);
                }
            }
        }
        if (!have_space || !have_layout || !is_i16) throw ParseError{"Signal is not an int16 dataset"};
        if (len > (1ull << 32)) throw ParseError{"implausible Signal length"};
        const uint64_t full_len = len;
        if (head > 0 && head < len) len = head;
        out->assign(len, 0);
        if (b_.u8(layout.off) != 3) throw ParseError{"unsupported layout version"};
        const int cls = b_.u8(layout.off + 1);
        if (cls == 0) {
            const uint32_t size = b_.u16(layout.off + 2);
            b_.need(layout.off + 4, size);
            std::memcpy(out->data(), b_.p + layout.off + 4, std::min<uint64_t>(size, len * 2));
        } else if (cls == 1) {
            const uint64_t a = b_.u64(layout.off + 2);
            if (a != kUndef) {
                b_.need(a, len * 2);
                std::memcpy(out->data(), b_.p + a, len * 2);
            }
        } else if (cls == 2) {
            const int ndims = b_.u8(layout.off + 2);
            if (ndims != 2) throw ParseError{"unexpected chunk rank"};
            const uint64_t btree = b_.u64(layout.off + 3);
            const uint32_t chunk_elems = b_.u32(layout.off + 11);
            if (btree != kUndef && len) walk_chunks(btree, chunk_elems, filters, out, 0);
        } else {
            throw ParseError{"unsupported layout class"};
        }
        return full_len;
    }

  private:
    void walk_group(uint64_t addr, uint64_t heap_data, std::vector<Link>* out, int depth) const {
        if (depth > 16) throw ParseError{"group B-tree too deep"};
        if (b_.sig(addr, "TREE")) {
            const int n = b_.u16(addr + 6);
            uint64_t p = addr + 24;
            for (int i = 0; i < n; ++i) {
                walk_group(b_.u64(p + 8), heap_data, out, depth + 1);
                p += 16;
            }
        } else if (b_.sig(addr, "SNOD")) {
            const int n = b_.u16(addr + 6);
            uint64_t p = addr + 8;
            for (int i = 0; i < n; ++i) {
                const uint64_t name_off = b_.u64(p), obj = b_.u64(p + 8);
                const uint64_t s = heap_data + name_off;
                b_.need(s, 1);
                const size_t ln = strnlen(reinterpret_cast<const char*>(b_.p + s), b_.n - s);
                out->push_back({std::string(reinterpret_cast<const char*>(b_.p + s), ln), obj});
                p += 40;
            }
        } else {
            throw ParseError{"bad group B-tree node"};
        }
    }

    bool parse_link(uint64_t p, uint64_t limit, Link* l, uint64_t* end) const {
        if (p + 3 > limit || b_.u8(p) != 1) return false;
        const int flags = b_.u8(p + 1);
        uint64_t q = p + 2;
        int type = 0;
        if (flags & 0x08) type = b_.u8(q++);
        if (flags & 0x04) q += 8;
        if (flags & 0x10) q += 1;
        const int ls = 1 << (flags & 3);
        const uint64_t nlen = b_.uvar(q, ls);
        q += ls;
        if (q + nlen > limit) return false;
        l->name.assign(reinterpret_cast<const char*>(b_.p + q), nlen);
        q += nlen;
        if (type == 0) {
            if (q + 8 > limit) return false;
            l->addr = b_.u64(q);
            q += 8;
        } else if (type == 1) {
            q += 2 + b_.u16(q);
            l->addr = kUndef;
        } else {
            return false;
        }
        *end = q;
        return type == 0;
    }

    void dense_links(uint64_t fheap, uint64_t name_index, std::vector<Link>* out) const {
        if (!b_.sig(fheap, "FRHP")) throw ParseError{"bad fractal heap"};
        uint64_t p = fheap + 5;
        const uint32_t io_filter_len = b_.u16(p + 2);
        const int flags = b_.u8(p + 4);
        p += 5;
        const uint32_t max_managed = b_.u32(p);
        p += 4 + 8 * 12;
        const uint32_t width = b_.u16(p);
        p += 2;
        const uint64_t start_size = b_.u64(p), max_direct = b_.u64(p + 8);
        p += 16;
        const uint32_t max_heap_bits = b_.u16(p);
        p += 4;
        const uint64_t root = b_.u64(p);
        const uint32_t cur_rows = b_.u16(p + 8);
        const int off_bytes = (max_heap_bits + 7) / 8;
        const bool checksummed = flags & 0x02;
        if (root == kUndef) return;
        struct Block { uint64_t heap_off, addr, size; };
        std::vector<Block> blocks;
        auto add_direct = [&](uint64_t a, uint64_t size) {
            if (a == kUndef || a + size > b_.n || !b_.sig(a, "FHDB")) return;
            blocks.push_back({b_.uvar(a + 13, off_bytes), a, size});
        };
        if (cur_rows == 0) {
            add_direct(root, start_size);
        } else {
            if (!b_.sig(root, "FHIB")) throw ParseError{"bad fractal heap indirect block"};
            uint64_t q = root + 5 + 8 + off_bytes;
            uint32_t max_direct_rows = 2;
            for (uint64_t s = start_size; s < max_direct; s *= 2) ++max_direct_rows;
            for (uint32_t row = 0; row < std::min(cur_rows, max_direct_rows); ++row) {
                const uint64_t row_size = start_size * (row < 2 ? 1 : (1ull << (row - 1)));
                for (uint32_t c = 0; c < width; ++c) {
                    add_direct(b_.u64(q), row_size);
                    q += 8 + (io_filter_len ? 12 : 0);
                }
            }
        }
        // live objects: heap ids from a single-leaf v2 B-tree name index
        if (name_index != kUndef && b_.sig(name_index, "BTHD") && b_.u8(name_index + 5) == 5 &&
            b_.u16(name_index + 12) == 0) {
            const uint32_t rec = b_.u16(name_index + 10);
            const uint64_t leaf = b_.u64(name_index + 16);
            const uint32_t nrec = b_.u16(name_index + 24);
            if (nrec == 0 || leaf == kUndef) return;
            if (b_.sig(leaf, "BTLF")) {
                auto bits_to_bytes = [](uint64_t v) { int n = 0; while (v) { ++n; v >>= 1; } return (n + 7) / 8; };
                const int len_bytes = std::min(bits_to_bytes(max_direct), bits_to_bytes(max_managed));
                for (uint32_t i = 0; i < nrec; ++i) {
                    const uint64_t id = leaf + 6 + static_cast<uint64_t>(i) * rec + 4;
                    if ((b_.u8(id) >> 4) & 3) continue;
                    const uint64_t off = b_.uvar(id + 1, off_bytes);
                    const uint64_t len = b_.uvar(id + 1 + off_bytes, len_bytes);
                    for (const Block& bl : blocks)
                        if (bl.heap_off <= off && off < bl.heap_off + bl.size) {
                            Link l;
                            uint64_t end;
                            const uint64_t q = bl.addr + (off - bl.heap_off);
                            if (parse_link(q, std::min<uint64_t>(q + len, b_.n), &l, &end)) out->push_back(l);
                            break;
                        }
                }
                return;
            }
        }
        // fallback: linear scan of the direct blocks (may include stale links left in free space)
        const uint64_t hdr = 5 + 8 + off_bytes + (checksummed ? 4 : 0);
        for (const Block& bl : blocks) {
            uint64_t q = bl.addr + hdr;
            const uint64_t end = bl.addr + bl.size;
            while (q + 10 < end) {
                Link l;
                uint64_t nq = q;
                if (!parse_link(q, end, &l, &nq) || nq <= q) break;
                out->push_back(l);
                q = nq;
            }
        }
    }

    void walk_chunks(uint64_t addr, uint32_t chunk_elems, const std::vector<int>& filters,
                     std::vector<int16_t>* out, int depth) const {
        if (depth > 16 || !b_.sig(addr, "TREE")) throw ParseError{"bad chunk B-tree"};
        const int level = b_.u8(addr + 5), n = b_.u16(addr + 6);
        const uint64_t key = 8 + 8 * 2;
        uint64_t p = addr + 24;
        for (int i = 0; i < n; ++i) {
            const uint32_t csize = b_.u32(p), mask = b_.u32(p + 4);
            const uint64_t off = b_.u64(p + 8);
            const uint64_t child = b_.u64(p + key);
            p += key + 8;
            if (level > 0) {
                walk_chunks(child, chunk_elems, filters, out, depth + 1);
                continue;
            }
            if (off >= out->size()) continue;   // beyond what is wanted of this data set
            b_.need(child, csize);
            // Common case - deflate only (every fast5 of the reference's fixtures): inflate
            // straight into the output signal, no intermediate copies (a chunk is declared as 201 536 elements
            // in these files whatever the read length, so sizing a scratch buffer by it costs more than the
            // inflate itself).
            int active = 0, only = -1;
            for (int f = 0; f < static_cast<int>(filters.size()); ++f)
                if (!(mask & (1u << f))) {
                    ++active;
                    only = filters[f];
                }
            if (active == 0) {
                if (off < out->size()) {
                    const uint64_t cnt = std::min<uint64_t>({csize / 2, chunk_elems, out->size() - off});
                    std::memcpy(out->data() + off, b_.p + child, cnt * 2);
                }
                continue;
            }
            if (active == 1 && only == 1) {
                if (off < out->size()) {
                    const uint64_t cnt = std::min<uint64_t>(chunk_elems, out->size() - off);
                    inflate_into(b_.p + child, csize, reinterpret_cast<uint8_t*>(out->data() + off), cnt * 2);
                }
                continue;
            }
            std::vector<uint8_t> raw(b_.p + child, b_.p + child + csize), tmp;
            for (int f = static_cast<int>(filters.size()) - 1; f >= 0; --f) {
                if (mask & (1u << f)) continue;
                if (filters[f] == 1) {
                    tmp.resize(static_cast<size_t>(chunk_elems) * 2 + 64);
                    const size_t got = zlib_inflate_into(raw.data(), raw.size(), tmp.data(), tmp.size());
                    tmp.resize(got);
                    raw.swap(tmp);
                } else if (filters[f] == 2) {
                    const size_t cnt = raw.size() / 2;
                    tmp.resize(cnt * 2);
                    for (size_t e = 0; e < cnt; ++e) {
                        tmp[2 * e] = raw[e];
                        tmp[2 * e + 1] = raw[cnt + e];
                    }
                    raw.swap(tmp);
                } else if (filters[f] == 3) {
                    if (raw.size() >= 4) raw.resize(raw.size() - 4);
                } else {
                    throw ParseError{"unsupported HDF5 filter (e.g. VBZ)"};
                }
            }
            // some writers store a truncated edge chunk: copy what exists
            if (off < out->size()) {
                const uint64_t cnt = std::min<uint64_t>({raw.size() / 2, chunk_elems, out->size() - off});
                std::memcpy(out->data() + off, raw.data(), cnt * 2);
            }
        }
    }

    Buf b_;
    uint64_t root_ = 0;
};

// Whole file into `data` with plain open / fstat / read (no stdio buffering: one system call per 1 MB).
bool read_file(const char* path, std::vector<uint8_t>* data, size_t* file_size) {
    const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return false;
    struct stat sb;
    if (::fstat(fd, &sb) != 0 || sb.st_size <= 0) {
        ::close(fd);
        return false;
    }
    const size_t size = static_cast<size_t>(sb.st_size);
    if (data->size() < size + kFilePadding) data->resize(size + kFilePadding);   // (the decoder may over-read a few bytes)
    *file_size = size;
    size_t got = 0;
    while (got < size) {
        const ssize_t r = ::read(fd, data->data() + got, size - got);
        if (r <= 0) break;
        got += static_cast<size_t>(r);
    }
    ::close(fd);
    return got == size;
}

struct ReadRec {
    std::string id;
    std::vector<int16_t> signal;
    uint64_t full_length = 0;
};

// One read group (`/Raw/Reads/Read_<n>` of the old layout, `/read_<uuid>/Raw` of the new one):
// `read_id` attribute + `Signal` dataset (load_fast5s.py:33-44).
bool read_group(const File& f, uint64_t group, ReadRec* r, uint64_t head) {
    if (!f.string_attr(group, "read_id", &r->id)) return false;
    uint64_t sig;
    if (!f.find(group, "Signal", &sig)) return false;
    r->full_length = f.read_i16(sig, &r->signal, head);
    return true;
}

// All reads of a fast5 file.  Single-read layouts give one read; a multi-read file (several
// `/read_<uuid>` groups in the root, load_fast5s.py:67-98) gives every read, ordered by group name -
// read straight from the file, where the reference unpacks it to one file per read with ONT's
// `multi_to_single_fast5` first (realtime.py:183-196).  A read that cannot be parsed is skipped, as
// an unreadable unpacked file would be (load_fast5s.py:48-49).
// status: 0 ok, 1 unreadable / not HDF5 / no read.  *multi = the root holds more than one read group.
int read_all(const char* path, std::vector<ReadRec>* reads, bool* multi, bool want_multi, uint64_t head = 0) {
    *multi = false;
    thread_local std::vector<uint8_t> data;   // reused: no allocation / zero fill per file once it has grown
    size_t file_size = 0;
    if (!read_file(path, &data, &file_size)) return 1;
    try {
        File f(data, file_size);
        const std::vector<Link> root = f.links(f.root());
        for (const Link& l : root)
            if (l.name == "Raw") {   // old single-read layout: /Raw/Reads/<first child>
                uint64_t rg;
                if (!f.find(l.addr, "Reads", &rg)) return 1;
                const std::vector<Link> kids = f.links(rg);
                if (kids.empty()) return 1;
                ReadRec r;
                if (!read_group(f, kids[0].addr, &r, head)) return 1;
                reads->push_back(std::move(r));
                return 0;
            }
        std::vector<const Link*> groups;   // new layout: /read_<uuid>/Raw
        for (const Link& l : root)
            if (l.name.compare(0, 5, "read_") == 0) groups.push_back(&l);
        if (groups.empty()) return 1;
        *multi = groups.size() > 1;
        if (*multi && !want_multi) return 0;
        std::sort(groups.begin(), groups.end(), [](const Link* a, const Link* b) { return a->name < b->name; });
        for (const Link* l : groups) {
            ReadRec r;
            uint64_t raw;
            try {
                if (f.find(l->addr, "Raw", &raw) && read_group(f, raw, &r, head)) reads->push_back(std::move(r));
            } catch (const ParseError&) {
                if (!*multi) return 1;
            }
        }
        return reads->empty() ? 1 : 0;
    } catch (const ParseError&) {
        return 1;
    } catch (const std::exception&) {
        return 1;
    }
}

// status: 0 ok, 1 unreadable / not HDF5 / no read (-> (None, None) in the reference), 2 multi-read file
int read_one(const char* path, std::string* read_id, std::vector<int16_t>* signal) {
    std::vector<ReadRec> reads;
    bool multi = false;
    const int st = read_all(path, &reads, &multi, false);
    if (st) return st;
    if (multi) return 2;
    *read_id = std::move(reads[0].id);
    *signal = std::move(reads[0].signal);
    return 0;
}

}  // namespace

struct db_fast5_batch {
    std::vector<int16_t> samples;
    std::vector<int64_t> offsets;      // rows + 1
    std::vector<int64_t> full_length;  // untruncated signal length per row
    std::vector<char> read_ids;        // rows x 64, NUL padded
    std::vector<int32_t> status;       // per row: 0 ok, 1 unreadable, 2 multi-read (rejected)
    std::vector<int32_t> row_file;     // per row: index of the file it came from
};

#pragma GCC visibility push(default)
extern "C" {

int db_zlib_inflate(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t dst_capacity, int64_t* produced, int use_zlib) {
    if (!src || !dst || !produced || src_len < 0 || dst_capacity < 0) return dbn::fail(DBN_EINVAL, "db_zlib_inflate: bad argument");
    std::vector<uint8_t> padded(static_cast<size_t>(src_len) + kFilePadding, 0);
    std::memcpy(padded.data(), src, static_cast<size_t>(src_len));
    try {
        size_t got = 0;
        if (use_zlib) got = zlib_inflate_into(padded.data(), static_cast<size_t>(src_len), dst, static_cast<size_t>(dst_capacity));
        else if (!dbn_inflate::inflate_zlib(padded.data(), static_cast<size_t>(src_len), dst, static_cast<size_t>(dst_capacity), &got)) return 1;
        *produced = static_cast<int64_t>(got);
        return 0;
    } catch (const ParseError&) {
        return 1;
    }
}

int db_fast5_read(const char* path, char* read_id, int16_t* signal, int64_t capacity, int64_t* length) {
    if (!path || !read_id || !length) return dbn::fail(DBN_EINVAL, "db_fast5_read: NULL argument");
    std::string id;
    std::vector<int16_t> sig;
    const int st = read_one(path, &id, &sig);
    if (st) return st;
    *length = static_cast<int64_t>(sig.size());
    std::memset(read_id, 0, 64);
    std::memcpy(read_id, id.data(), std::min<size_t>(id.size(), 63));
    if (signal && capacity >= *length) std::memcpy(signal, sig.data(), sig.size() * 2);
    return 0;
}

int db_fast5_list_root(const char* path, char* names, int64_t capacity, int* count) {
    if (!path || !names || !count) return dbn::fail(DBN_EINVAL, "db_fast5_list_root: NULL argument");
    *count = 0;
    std::vector<uint8_t> data;
    size_t file_size = 0;
    if (!read_file(path, &data, &file_size)) return 1;
    try {
        File f(data, file_size);
        int64_t used = 0;
        for (const Link& l : f.links(f.root())) {
            const int64_t need = static_cast<int64_t>(l.name.size()) + 1;
            if (used + need > capacity) break;
            std::memcpy(names + used, l.name.c_str(), need);
            used += need;
            ++*count;
        }
        return 0;
    } catch (const ParseError&) {
        return 1;
    } catch (const std::exception&) {
        return 1;
    }
}

// multi = 0: one row per file (a multi-read file is a row with status 2);
// multi = 1: one row per READ - a multi-read file contributes all its reads, an unreadable file one
// row with status 1.  Rows are in file order (reads of a file ordered by group name).
// sides: bit 0 = the start of the reads is needed, bit 1 = the end (start only: reads are cut after `keep`
// samples and the inflate stops there).
static int batch_read(const char* const* paths, int n, int threads, int64_t keep, int multi, int sides, db_fast5_batch** out) {
    if (!paths || n < 0 || !out) return dbn::fail(DBN_EINVAL, "db_fast5_batch_read: bad argument");
    db_fast5_batch* b = new db_fast5_batch();
    std::vector<std::vector<ReadRec>> per_file(n);
    std::vector<int32_t> file_status(n, 1);
    std::vector<std::vector<int64_t>> full(n);
    std::atomic<int> next{0};
    auto work = [&]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
            bool is_multi = false;
            const uint64_t head = (keep > 0 && sides == 1) ? static_cast<uint64_t>(keep) : 0;
            file_status[i] = read_all(paths[i], &per_file[i], &is_multi, multi != 0, head);
            if (file_status[i] == 0 && is_multi && !multi) file_status[i] = 2;
            if (file_status[i]) {
                per_file[i].clear();
                continue;
            }
            for (ReadRec& r : per_file[i]) {
                full[i].push_back(static_cast<int64_t>(r.full_length));
                // keep only what call_batch can ever look at: the first and last `keep` samples
                if (keep > 0 && static_cast<int64_t>(r.signal.size()) > 2 * keep) {
                    r.signal.erase(r.signal.begin() + keep, r.signal.end() - keep);
                    r.signal.shrink_to_fit();
                }
            }
        }
    };
    const int nt = std::max(1, std::min(threads, n));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
    b->offsets.assign(1, 0);
    for (int i = 0; i < n; ++i) {
        if (file_status[i]) {   // one placeholder row so that every file is represented
            b->status.push_back(file_status[i]);
            b->row_file.push_back(i);
            b->full_length.push_back(0);
            b->offsets.push_back(b->offsets.back());
            continue;
        }
        for (size_t k = 0; k < per_file[i].size(); ++k) {
            b->status.push_back(0);
            b->row_file.push_back(i);
            b->full_length.push_back(full[i][k]);
            b->offsets.push_back(b->offsets.back() + static_cast<int64_t>(per_file[i][k].signal.size()));
        }
    }
    const size_t rows = b->status.size();
    b->samples.resize(std::max<int64_t>(b->offsets.back(), 1));
    b->read_ids.assign(rows * 64, 0);
    size_t row = 0;
    for (int i = 0; i < n; ++i) {
        if (file_status[i]) {
            ++row;
            continue;
        }
        for (const ReadRec& r : per_file[i]) {
            std::memcpy(b->samples.data() + b->offsets[row], r.signal.data(), r.signal.size() * 2);
            std::memcpy(b->read_ids.data() + row * 64, r.id.data(), std::min<size_t>(r.id.size(), 63));
            ++row;
        }
    }
    *out = b;
    return 0;
}

int db_fast5_batch_read(const char* const* paths, int n, int threads, int64_t keep, db_fast5_batch** out) {
    return batch_read(paths, n, threads, keep, 0, 3, out);
}

int db_fast5_batch_read_reads(const char* const* paths, int n, int threads, int64_t keep, db_fast5_batch** out) {
    return batch_read(paths, n, threads, keep, 1, 3, out);
}

int db_fast5_batch_read_sides(const char* const* paths, int n, int threads, int64_t keep, int sides, db_fast5_batch** out) {
    if (sides < 1 || sides > 3) return dbn::fail(DBN_EINVAL, "db_fast5_batch_read_sides: sides must be 1 (start), 2 (end) or 3");
    return batch_read(paths, n, threads, keep, 1, sides, out);
}

int db_fast5_batch_rows(const db_fast5_batch* b, int64_t* rows, const int32_t** row_file) {
    if (!b || !rows) return dbn::fail(DBN_EINVAL, "db_fast5_batch_rows: NULL argument");
    *rows = static_cast<int64_t>(b->status.size());
    if (row_file) *row_file = b->row_file.data();
    return 0;
}

int db_fast5_batch_get(const db_fast5_batch* b, const int16_t** samples, const int64_t** offsets,
                       const int64_t** full_length, const char** read_ids, const int32_t** status) {
    if (!b) return dbn::fail(DBN_EINVAL, "db_fast5_batch_get: NULL batch");
    if (samples) *samples = b->samples.data();
    if (offsets) *offsets = b->offsets.data();
    if (full_length) *full_length = b->full_length.data();
    if (read_ids) *read_ids = b->read_ids.data();
    if (status) *status = b->status.data();
    return 0;
}

void db_fast5_batch_free(db_fast5_batch* b) { delete b; }

}  // extern "C"
#pragma GCC visibility pop
