// pf_server_check.cpp — compile-and-run check of the C++ host mirror (prefhetch::Server) against
// the C ABI: plaintext stages on a tiny synthetic index.  Built by __graft_entry__.build();
// executed by tests/test_gpu_parity.py::test_cpp_host_mirror on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pf_server.hpp"

int main() {
    const uint32_t d = 128;
    const uint64_t nlist = 4, per = 50, nb = nlist * per;
    std::vector<uint64_t> primes = {0x7fffffd8001ULL, 0x7fffffc8001ULL, 0xfffffffc001ULL, 0xffffff6c001ULL,
                                    0xfffffebc001ULL};
    try {
        prefhetch::Server srv(d, 8192, primes, 16760833);
        std::vector<float> cent(nlist * d), vec(nb * d);
        std::vector<prefhetch::idx_t> off(nlist + 1), ids(nb);
        unsigned s = 12345;
        auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 16) & 0xff; };
        for (auto &c : cent) c = static_cast<float>(rnd());
        for (auto &v : vec) v = static_cast<float>(rnd());
        for (uint64_t l = 0; l <= nlist; l++) off[l] = static_cast<prefhetch::idx_t>(l * per);
        for (uint64_t i = 0; i < nb; i++) ids[i] = static_cast<prefhetch::idx_t>(i);
        srv.init_index(nlist, cent.data(), off.data(), ids.data(), vec.data());
        std::vector<float> q(vec.begin(), vec.begin() + 2 * d), back;
        srv.retrieve_centroids(back);
        if (back != cent) return 2;
        std::vector<prefhetch::idx_t> probes;
        srv.coarseQuantize(q, 2, probes);
        std::vector<float> dist;
        std::vector<prefhetch::idx_t> labels;
        std::vector<size_t> sizes;
        srv.coarseSearch(q, probes, 2, dist, labels, sizes);
        if (sizes.size() != 2 || sizes[0] != 2 * per || dist.size() != 4 * per) return 3;
        // query 0 is base vector 0: if its list is probed its distance must be exactly 0
        bool found = false;
        for (size_t i = 0; i < sizes[0]; i++)
            if (labels[i] == 0) found = dist[i] == 0.0f;
        std::vector<prefhetch::idx_t> rows = {0, 1, 2, 3};
        std::vector<float> pd;
        srv.preciseSearch(std::span<const float>(q.data(), d), rows, 4, pd);
        if (pd[0] != 0.0f) return 4;
        std::printf("pf_server_check ok (list of vector 0 probed: %d)\n", found ? 1 : 0);
        try {
            std::vector<prefhetch::idx_t> bad = {99, 0, 1, 2};
            srv.coarseSearch(q, bad, 2, dist, labels, sizes);
            return 5; // must have thrown like the reference does
        } catch (const std::runtime_error &) {
        }
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "pf_server_check failed: %s\n", ex.what());
        return 1;
    }
    return 0;
}
