// pf_faiss_check.cpp — CPU-only check of pf_faiss_io.hpp: reads the file named on the command line
// and prints its shape plus FNV-1a checksums, compared by tests/test_faiss_io.py with the Python
// reader.  Built by __graft_entry__.build().
#include <cstdio>

#include "pf_faiss_io.hpp"

static uint64_t fnv(const void *p, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) h = (h ^ static_cast<const uint8_t *>(p)[i]) * 1099511628211ULL;
    return h;
}

int main(int argc, char **argv) {
    if (argc != 2) return 2;
    try {
        const prefhetch::IvfFile f = prefhetch::read_ivfpq_file(argv[1]);
        std::printf("%u %llu %llu %llu %llu %016llx %016llx %016llx %llu %llu %016llx %016llx\n", f.d, (unsigned long long)f.ntotal,
                    (unsigned long long)f.nlist, (unsigned long long)f.nprobe, (unsigned long long)f.code_size,
                    (unsigned long long)fnv(f.centroids.data(), f.centroids.size() * 4),
                    (unsigned long long)fnv(f.list_offsets.data(), f.list_offsets.size() * 8),
                    (unsigned long long)fnv(f.ids.data(), f.ids.size() * 8), (unsigned long long)f.pq_M,
                    (unsigned long long)f.pq_nbits, (unsigned long long)fnv(f.pq_centroids.data(), f.pq_centroids.size() * 4),
                    (unsigned long long)fnv(f.codes.data(), f.codes.size()));
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "%s\n", ex.what());
        return 1;
    }
    return 0;
}
