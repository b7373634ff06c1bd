// pf_server_check.cpp — compile-and-run check of the C++ host mirror (prefhetch::Server) against
// the C ABI.  Built by __graft_entry__.build(); executed on the GPU box by tests/test_gpu_parity.py:
//   pf_server_check              plaintext stages on a tiny synthetic index (test_cpp_host_mirror)
//   pf_server_check <dir>        the ENCRYPTED search from C++: index, SEAL GaloisKeys stream, serialized
//                                query ciphertexts and probe lists are read from files the test wrote;
//                                the result ciphertext streams are written to <dir>/results.bin, one after
//                                the other, for a byte comparison with the Python path (and the oracle)
//                                (test_cpp_encrypted_search_matches_python).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "pf_query_handlers.hpp"
#include "pf_server.hpp"

namespace {

template <class T>
std::vector<T> read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<T> v(static_cast<size_t>(n) / sizeof(T));
    f.read(reinterpret_cast<char *>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
    return v;
}

int encrypted_check(const std::string &dir) {
    // params.txt: d n g m result_limbs nprobe nq t k prime_0 .. prime_{k-1}
    std::ifstream pf(dir + "/params.txt");
    if (!pf) throw std::runtime_error("cannot open params.txt");
    uint32_t d, g, m, rl, nprobe;
    uint64_t n, nq, t, k;
    pf >> d >> n >> g >> m >> rl >> nprobe >> nq >> t >> k;
    std::vector<uint64_t> primes(k);
    for (auto &q : primes) pf >> q;
    const auto cent = read_file<float>(dir + "/centroids.f32");
    const auto off = read_file<prefhetch::idx_t>(dir + "/offsets.i64");
    const auto ids = read_file<prefhetch::idx_t>(dir + "/ids.i64");
    const auto vec = read_file<float>(dir + "/vectors.f32");
    const auto keys = read_file<uint8_t>(dir + "/galois_keys.bin");
    const auto qblob = read_file<uint8_t>(dir + "/queries.bin");
    const auto qoff = read_file<uint64_t>(dir + "/query_offsets.u64");
    const auto probes = read_file<prefhetch::idx_t>(dir + "/probes.i64");
    const uint64_t nlist = off.size() - 1;
    prefhetch::Server srv(d, n, primes, t, m, g, 0, 0, 1, rl);
    srv.init_index(nlist, cent.data(), off.data(), ids.data(), vec.data());
    srv.loadGaloisKeys(keys);
    // the synchronous call, then the same request twice through submit / collect: all three must agree
    prefhetch::EncryptedCoarseResult r0, r1, r2;
    srv.coarseSearchEncrypted(nq, qblob, qoff, probes, nprobe, r0);
    const uint64_t t1 = srv.submitSearchEncrypted(nq, qblob, qoff, probes, nprobe, r1);
    const uint64_t t2 = srv.submitSearchEncrypted(nq, qblob, qoff, probes, nprobe, r2);
    srv.collect(t1, r1);
    srv.collect(t2, r2);
    if (r0.ciphertexts != r1.ciphertexts || r0.ciphertexts != r2.ciphertexts || r0.result_offsets != r2.result_offsets ||
        r0.coarse_vector_indexes != r1.coarse_vector_indexes)
        return 6;
    const size_t len = srv.resultSerializedSize();
    std::ofstream out(dir + "/results.bin", std::ios::binary);
    for (uint64_t r = 0; r < r0.nresults; r++)
        out.write(reinterpret_cast<const char *>(r0.ciphertexts.data() + r0.result_offsets[r]), static_cast<std::streamsize>(len));
    std::ofstream lab(dir + "/labels.i64", std::ios::binary);
    lab.write(reinterpret_cast<const char *>(r0.coarse_vector_indexes.data()),
              static_cast<std::streamsize>(r0.coarse_vector_indexes.size() * sizeof(prefhetch::idx_t)));
    std::printf("pf_server_check encrypted ok: %llu results of %zu bytes, %zu labels\n", (unsigned long long)r0.nresults, len,
                r0.coarse_vector_indexes.size());
    // a malformed request is an exception (std::runtime_error, as the reference throws), not a crash
    try {
        std::vector<uint64_t> bad(qoff);
        bad.back() += 64;
        srv.coarseSearchEncrypted(nq, qblob, bad, probes, nprobe, r1);
        return 7;
    } catch (const std::runtime_error &) {
    }
    return 0;
}

// `--handlers <dir>`: the handler bodies of pf_query_handlers.hpp end to end — JSON request bodies in the
// reference's shape in, JSON responses out — against direct calls on the same Server (files as in encrypted_check).
int handlers_check(const std::string &dir) {
    namespace h = prefhetch::handlers;
    std::ifstream pf(dir + "/params.txt");
    if (!pf) throw std::runtime_error("cannot open params.txt");
    uint32_t d, g, m, rl, nprobe;
    uint64_t n, nq, t, k;
    pf >> d >> n >> g >> m >> rl >> nprobe >> nq >> t >> k;
    std::vector<uint64_t> primes(k);
    for (auto &q : primes) pf >> q;
    const auto cent = read_file<float>(dir + "/centroids.f32");
    const auto off = read_file<prefhetch::idx_t>(dir + "/offsets.i64");
    const auto ids = read_file<prefhetch::idx_t>(dir + "/ids.i64");
    const auto vec = read_file<float>(dir + "/vectors.f32");
    const auto keys = read_file<uint8_t>(dir + "/galois_keys.bin");
    const auto qblob = read_file<uint8_t>(dir + "/queries.bin");
    const auto qoff = read_file<uint64_t>(dir + "/query_offsets.u64");
    const auto probes = read_file<prefhetch::idx_t>(dir + "/probes.i64");
    const auto queries = read_file<float>(dir + "/queries.f32");
    const uint64_t nlist = off.size() - 1;
    prefhetch::Server srv(d, n, primes, t, m, g, 0, 0, 1, rl);
    srv.init_index(nlist, cent.data(), off.data(), ids.data(), vec.data());
    srv.loadGaloisKeys(keys);
    // GET /query
    size_t r = 0, c = 0;
    const auto cent2 = h::json::matrix<float>(h::query(srv), r, c);
    if (r != nlist || c != d || cent2 != cent) return 10;
    // POST /coarsesearch
    std::string body = "{\"preciseQuery\":";
    h::json::put_matrix(body, queries.data(), nq, d);
    body += ",\"nearestCentroidIndexes\":";
    h::json::put_matrix(body, probes.data(), nq, nprobe);
    body += "}";
    const std::string resp_body = h::coarse_search(srv, body); // the parsed views point into it
    const auto resp = h::json::object(resp_body);
    std::vector<float> dist;
    std::vector<prefhetch::idx_t> labels;
    std::vector<size_t> sizes;
    srv.coarseSearch(queries, probes, nprobe, dist, labels, sizes);
    if (h::json::vector<float>(h::json::at(resp, "coarseDistanceScores")) != dist) return 11;
    if (h::json::vector<prefhetch::idx_t>(h::json::at(resp, "coarseVectorIndexes")) != labels) return 12;
    if (h::json::vector<size_t>(h::json::at(resp, "listSizesPerQuery")) != sizes) return 13;
    // POST /precisesearch on the first candidates of every query
    const size_t probe = 4;
    std::vector<prefhetch::idx_t> cand(nq * probe);
    size_t o = 0;
    for (uint64_t i = 0; i < nq; i++) {
        for (size_t j = 0; j < probe; j++) cand[i * probe + j] = labels[o + j];
        o += sizes[i];
    }
    std::string pbody = "{\"preciseQuery\":";
    h::json::put_matrix(pbody, queries.data(), nq, d);
    pbody += ",\"nearestCoarseVectorIndexes\":";
    h::json::put_matrix(pbody, cand.data(), nq, probe);
    pbody += "}";
    const std::string presp_body = h::precise_search(srv, pbody);
    const auto presp = h::json::object(presp_body);
    std::vector<float> pd;
    srv.preciseSearch(queries, cand, probe, pd);
    if (h::json::matrix<float>(h::json::at(presp, "preciseDistanceScores"), r, c) != pd || r != nq || c != probe) return 14;
    o = 0;
    for (uint64_t i = 0; i < nq; i++) { // plaintext stage 2 and the exact re-rank agree on integer data
        for (size_t j = 0; j < probe; j++)
            if (pd[i * probe + j] != dist[o + j]) return 15;
        o += sizes[i];
    }
    // POST /galoiskeys: the same key stream through the handler (replaces the set loaded above with itself)
    {
        const std::string kbody = "{\"galoisKeys\":\"" + h::json::base64_encode(keys.data(), keys.size()) + "\"}";
        const std::string kresp_body = h::galois_keys(srv, kbody);
        const auto kresp = h::json::object(kresp_body);
        if (h::json::number<uint64_t>(h::json::at(kresp, "galoisKeysBytes")) != keys.size()) return 21;
    }
    // POST /coarsesearch-encrypted
    std::string ebody = "{\"queryCiphertexts\":\"" + h::json::base64_encode(qblob.data(), qblob.size()) + "\",\"ctOffsets\":";
    h::json::put_vector(ebody, qoff.data(), qoff.size());
    ebody += ",\"nearestCentroidIndexes\":";
    h::json::put_matrix(ebody, probes.data(), nq, nprobe);
    ebody += "}";
    const std::string eresp_body = h::coarse_search_encrypted(srv, ebody);
    const auto eresp = h::json::object(eresp_body);
    prefhetch::EncryptedCoarseResult direct;
    srv.coarseSearchEncrypted(nq, qblob, qoff, probes, nprobe, direct);
    if (h::json::base64_decode(h::json::string(h::json::at(eresp, "resultCiphertexts"))) != direct.ciphertexts) return 16;
    if (h::json::vector<uint64_t>(h::json::at(eresp, "resultOffsets")) != direct.result_offsets) return 17;
    if (h::json::vector<prefhetch::idx_t>(h::json::at(eresp, "coarseVectorIndexes")) != direct.coarse_vector_indexes) return 18;
    if (h::json::number<uint64_t>(h::json::at(eresp, "resultBytes")) != srv.resultSerializedSize()) return 19;
    size_t pr = 0, pc = 0;
    if (h::json::matrix<uint64_t>(h::json::at(eresp, "probedListSizes"), pr, pc) != direct.probed_sizes || pr != nq || pc != nprobe) return 20;
    return 0;
}

} // namespace

int main(int argc, char **argv) {
    if (argc > 2 && std::string(argv[1]) == "--handlers") {
        try {
            const int rc = handlers_check(argv[2]);
            if (rc == 0) std::printf("pf_server_check handlers ok\n");
            return rc;
        } catch (const std::exception &ex) {
            std::fprintf(stderr, "pf_server_check (handlers) failed: %s\n", ex.what());
            return 1;
        }
    }
    if (argc > 1) {
        try {
            return encrypted_check(argv[1]);
        } catch (const std::exception &ex) {
            std::fprintf(stderr, "pf_server_check (encrypted) failed: %s\n", ex.what());
            return 1;
        }
    }

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
