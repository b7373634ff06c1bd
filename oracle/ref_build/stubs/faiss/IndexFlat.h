// Compile-only stand-in for the FAISS fork (absent; un-vendored FetchContent of PES-Innovation-Lab/PreFHEtch-faiss).
// train / add / write_index do nothing, so Server::init_index runs its "no cached index" branch up to and
// including the load of the base vectors, which is all Server::preciseSearch needs.  search_encrypted — the one
// call whose source is absent — aborts.
#pragma once
#include <cstdint>
#include <cstdlib>
namespace faiss {
using idx_t = int64_t;
struct Index {
    virtual ~Index() = default;
    virtual void reconstruct(idx_t, float *) const { std::abort(); }
};
struct IndexFlatL2 : Index {
    explicit IndexFlatL2(int64_t) {}
};
struct IndexIVFPQ : Index {
    IndexIVFPQ(Index *q, size_t, size_t, size_t, size_t) : quantizer(q) {}
    Index *quantizer;
    size_t nprobe = 1;
    void train(idx_t, const float *) {}
    void add(idx_t, const float *) {}
    void search_encrypted(idx_t, const float *, idx_t *, float *, idx_t *, size_t *) const { std::abort(); }
};
inline Index *read_index(const char *) { std::abort(); }
inline void write_index(const Index *, const char *) {}
} // namespace faiss
