// pf_faiss_io.hpp — reader for the FAISS IndexIVFPQ file the reference server caches on disk
// (ref: src/server/server_lib.cpp:38-42 builds the file name, :82 writes it, :91-95 reads it and
// rejects anything that is not an IndexIVFPQ).  What the GPU engine consumes is kept: the coarse
// centroids, the ids of every inverted list, and the product quantizer with the lists' codes (for the
// PQ-ADC distance of today's search_encrypted: pf_load_pq).
// [EXT, UNVERIFIED] layout restated from the published faiss/impl/index_read.cpp (SURVEY.md
// App. B.3); FAISS itself is not available here, so it is only checked against the writer in
// prefhetch_b200/faiss_io.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace prefhetch {

struct IvfFile {
    uint32_t d = 0;
    uint64_t ntotal = 0, nlist = 0, nprobe = 0, code_size = 0;
    std::vector<float> centroids;      // [nlist][d]
    std::vector<int64_t> list_offsets; // [nlist+1]
    std::vector<int64_t> ids;          // list order
    uint64_t pq_M = 0, pq_nbits = 0;   // ProductQuantizer: M sub-quantizers of nbits bits
    std::vector<float> pq_centroids;   // [M][2^nbits][d/M]
    std::vector<uint8_t> codes;        // [ntotal][code_size], list order
};

namespace detail {
class Cursor {
  public:
    explicit Cursor(const std::string &path) {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        if (!f) throw std::runtime_error("cannot open " + path);
        m_Buf.resize(static_cast<size_t>(f.tellg()));
        f.seekg(0);
        f.read(reinterpret_cast<char *>(m_Buf.data()), static_cast<std::streamsize>(m_Buf.size()));
    }
    template <class T> T get() {
        T v;
        copy(&v, sizeof(T));
        return v;
    }
    void copy(void *dst, size_t n) {
        if (n > m_Buf.size() - m_Pos) throw std::runtime_error("truncated index file");
        if (dst) std::memcpy(dst, m_Buf.data() + m_Pos, n);
        m_Pos += n;
    }
    void skip(size_t n) { copy(nullptr, n); }
    void skip_items(uint64_t count, uint64_t item_bytes) { // count * item_bytes without overflow
        if (item_bytes && count > (m_Buf.size() - m_Pos) / item_bytes) throw std::runtime_error("truncated index file");
        skip(static_cast<size_t>(count * item_bytes));
    }
    bool at_end() const { return m_Pos == m_Buf.size(); }
    size_t remaining() const { return m_Buf.size() - m_Pos; }

  private:
    std::vector<uint8_t> m_Buf;
    size_t m_Pos = 0;
};
constexpr uint32_t fourcc(const char (&s)[5]) {
    return uint32_t(uint8_t(s[0])) | uint32_t(uint8_t(s[1])) << 8 | uint32_t(uint8_t(s[2])) << 16 |
           uint32_t(uint8_t(s[3])) << 24;
}
inline void index_header(Cursor &c, uint32_t &d, uint64_t &ntotal) {
    d = static_cast<uint32_t>(c.get<int32_t>());
    ntotal = static_cast<uint64_t>(c.get<int64_t>());
    c.skip(16);
    c.skip(1); // is_trained
    if (c.get<int32_t>() > 1) c.skip(4); // metric_arg
}
} // namespace detail

inline IvfFile read_ivfpq_file(const std::string &path) {
    using namespace detail;
    Cursor c(path);
    IvfFile out;
    // "IwQR" = IndexIVFPQR, a subclass the reference's dynamic_cast<IndexIVFPQ*> accepts (its refine section after
    // the inverted lists is not needed here); "IvPQ" / "IvQR" = the legacy per-list layout, refused by name
    const uint32_t top = c.get<uint32_t>();
    if (top == fourcc("IvPQ") || top == fourcc("IvQR"))
        throw std::runtime_error("legacy IndexIVFPQ file layout (IvPQ / IvQR): re-save it with a current FAISS");
    if (top != fourcc("IwPQ") && top != fourcc("IwQR")) throw std::runtime_error("Loaded index is not of type IndexIVFPQ");
    index_header(c, out.d, out.ntotal);
    out.nlist = c.get<uint64_t>();
    out.nprobe = c.get<uint64_t>();
    if (out.d == 0 || out.nlist == 0 || out.nlist > (1ull << 32)) throw std::runtime_error("implausible index header");
    const uint32_t qcc = c.get<uint32_t>();
    if (qcc != fourcc("IxF2") && qcc != fourcc("IxFI") && qcc != fourcc("IxFl"))
        throw std::runtime_error("coarse quantizer is not an IndexFlat");
    uint32_t qd;
    uint64_t qn;
    index_header(c, qd, qn);
    const uint64_t nfloats = c.get<uint64_t>();
    if (qd != out.d || qn != out.nlist || nfloats != out.nlist * out.d)
        throw std::runtime_error("quantizer shape does not match the IVF header");
    if (nfloats > (1ull << 40)) throw std::runtime_error("truncated index file");
    out.centroids.resize(nfloats);
    c.copy(out.centroids.data(), nfloats * sizeof(float));
    const uint8_t dm = c.get<uint8_t>(); // DirectMap::NoMap / Array / Hashtable
    if (dm > 2) throw std::runtime_error("unknown direct-map type");
    c.skip_items(c.get<uint64_t>(), 8);
    if (dm == 2) c.skip_items(c.get<uint64_t>(), 16);
    c.skip(1); // by_residual
    out.code_size = c.get<uint64_t>();
    const uint64_t pq_d = c.get<uint64_t>();
    out.pq_M = c.get<uint64_t>();
    out.pq_nbits = c.get<uint64_t>();
    const uint64_t npq = c.get<uint64_t>();
    if (pq_d != out.d) throw std::runtime_error("product quantizer dimension does not match the index");
    if (npq > c.remaining() / sizeof(float)) throw std::runtime_error("truncated index file"); // before allocating
    out.pq_centroids.resize(npq);
    c.copy(out.pq_centroids.data(), npq * sizeof(float));
    const uint32_t il = c.get<uint32_t>();
    if (il == fourcc("il00")) throw std::runtime_error("index file holds no inverted lists (il00)");
    if (il != fourcc("ilar")) throw std::runtime_error("unsupported inverted-list container (only in-memory ArrayInvertedLists, 'ilar')");
    if (c.get<uint64_t>() != out.nlist || c.get<uint64_t>() != out.code_size)
        throw std::runtime_error("inverted lists do not match the index header");
    std::vector<uint64_t> sizes(out.nlist, 0);
    const uint32_t kind = c.get<uint32_t>();
    const uint64_t nsz = c.get<uint64_t>();
    if (kind == fourcc("full")) {
        if (nsz != out.nlist) throw std::runtime_error("bad list size table");
        c.copy(sizes.data(), nsz * 8);
    } else if (kind == fourcc("sprs")) {
        if (nsz % 2) throw std::runtime_error("bad list size table");
        for (uint64_t i = 0; i < nsz / 2; i++) {
            const uint64_t l = c.get<uint64_t>(), n = c.get<uint64_t>();
            if (l >= out.nlist) throw std::runtime_error("bad list size table");
            sizes[l] = n;
        }
    } else {
        throw std::runtime_error("unknown list size encoding");
    }
    out.list_offsets.assign(out.nlist + 1, 0);
    for (uint64_t l = 0; l < out.nlist; l++) {
        if (sizes[l] > out.ntotal) throw std::runtime_error("ntotal does not match the inverted lists");
        out.list_offsets[l + 1] = out.list_offsets[l] + static_cast<int64_t>(sizes[l]);
    }
    if (static_cast<uint64_t>(out.list_offsets[out.nlist]) != out.ntotal)
        throw std::runtime_error("ntotal does not match the inverted lists");
    // the codes and ids must be in the file before anything is allocated for them
    if (out.code_size > (1u << 20) || out.ntotal > c.remaining() / (out.code_size + 8)) throw std::runtime_error("truncated index file");
    out.ids.resize(out.ntotal);
    out.codes.resize(out.ntotal * out.code_size);
    for (uint64_t l = 0; l < out.nlist; l++) {
        c.copy(out.codes.data() + static_cast<uint64_t>(out.list_offsets[l]) * out.code_size, sizes[l] * out.code_size);
        c.copy(out.ids.data() + out.list_offsets[l], sizes[l] * 8);
    }
    if (top == fourcc("IwPQ") && !c.at_end()) throw std::runtime_error("trailing bytes after the inverted lists");
    return out;
}

} // namespace prefhetch
