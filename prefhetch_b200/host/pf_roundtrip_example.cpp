// pf_roundtrip_example.cpp — the reference's client main (ref: src/client/client.cpp:7-79) with the coarse step
// ENCRYPTED, client and server in one process and the HTTP hops replaced by the handler bodies' JSON strings:
//   get_centroids -> sort_nearest_centroids -> [keys, encrypted coarse query] -> POST /coarsesearch-encrypted ->
//   decrypt -> compute_nearest_coarse_vectors -> POST /precisesearch -> compute_nearest_precise_vectors -> recall.
// prefhetch::Client (pf_client.hpp, CPU) on one side, prefhetch::Server over the C ABI (pf_server.hpp, the GPU
// engine) behind prefhetch::handlers on the other; no SEAL, no oracle, no Python.  Synthetic SIFT-shaped data
// (deterministic).  Checks: the decrypted coarse scores equal the plaintext endpoint's scores (exact integers),
// the final ranking equals exact brute force over the probed lists.  Exit code 0 = all checks passed.
//   pf_roundtrip_example [nbase=6000] [nlist=48] [nquery=5] [nprobe=6]
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "pf_client.hpp"
#include "pf_query_handlers.hpp"

namespace {
uint64_t mix(uint64_t a) {
    a = (a ^ (a >> 31)) * 0x9E3779B97F4A7C15ULL;
    a = (a ^ (a >> 29)) * 0xBF58476D1CE4E5B9ULL;
    return a ^ (a >> 32);
}
std::string b64(const std::vector<uint8_t> &v) { return prefhetch::handlers::json::base64_encode(v.data(), v.size()); }
} // namespace

int main(int argc, char **argv) {
    namespace h = prefhetch::handlers;
    using prefhetch::idx_t;
    const uint64_t nb = argc > 1 ? strtoull(argv[1], nullptr, 10) : 6000, nlist = argc > 2 ? strtoull(argv[2], nullptr, 10) : 48;
    const uint64_t nq = argc > 3 ? strtoull(argv[3], nullptr, 10) : 5, nprobe = argc > 4 ? strtoull(argv[4], nullptr, 10) : 6;
    const uint32_t d = 128, m = 1, g = 8;
    const uint64_t N = 8192, t = 16760833; // BFVDefault(8192), 24-bit batching prime
    const std::vector<uint64_t> primes = {8796092858369ULL, 8796092792833ULL, 17592186028033ULL, 17592185438209ULL, 17592184717313ULL};
    const uint64_t coarse_probe = 40, K = 10;
    try {
        // ---- data: a clustered mixture of integer vectors, lists = nearest centre (what IndexIVF::add does) ----
        std::vector<float> centres(nlist * d), base(nb * d), query(nq * d);
        for (size_t i = 0; i < centres.size(); i++) centres[i] = (float)(mix(i + 11) % 200);
        auto draw = [&](uint64_t id, uint64_t salt, float *out) {
            const uint64_t c = mix(id * 7 + salt) % nlist;
            for (uint32_t k = 0; k < d; k++) {
                const int noise = (int)(mix(id * 131 + k + salt) % 41) - 20;
                out[k] = (float)std::min(255, std::max(0, (int)centres[c * d + k] + noise));
            }
        };
        for (uint64_t i = 0; i < nb; i++) draw(i, 1, base.data() + i * d);
        for (uint64_t i = 0; i < nq; i++) draw(i, 99991, query.data() + i * d);
        auto near_c = prefhetch::Client::sort_nearest_centroids(base.data(), nb, centres.data(), nlist, d);
        std::vector<std::vector<idx_t>> lists(nlist);
        for (uint64_t i = 0; i < nb; i++) lists[(size_t)near_c[i][0].idx].push_back((idx_t)i);
        std::vector<idx_t> offsets(nlist + 1, 0), ids;
        std::vector<float> vectors;
        for (uint64_t l = 0; l < nlist; l++) {
            offsets[l + 1] = offsets[l] + (idx_t)lists[l].size();
            for (idx_t id : lists[l]) {
                ids.push_back(id);
                vectors.insert(vectors.end(), base.begin() + id * d, base.begin() + (id + 1) * d);
            }
        }
        // ---- server (GPU) and client (CPU) ----
        prefhetch::Server srv(d, N, primes, t, m, g, /*device*/ 0, 0, 1, /*result_limbs*/ 1);
        srv.init_index(nlist, centres.data(), offsets.data(), ids.data(), vectors.data());
        prefhetch::Client he(d, N, primes, t, m, g);
        std::array<uint8_t, 64> seed;
        for (int i = 0; i < 64; i++) seed[i] = (uint8_t)mix(i + 5);
        he.generateKeys(seed);
        // POST /galoiskeys
        h::galois_keys(srv, "{\"galoisKeys\":\"" + b64(he.galoisKeys()) + "\"}");
        // GET /query -> centroids; stage 1 on the client, as in the reference
        size_t r = 0, c = 0;
        const std::vector<float> cent = h::json::matrix<float>(h::query(srv), r, c);
        if (r != nlist || c != d) throw std::runtime_error("GET /query shape");
        const auto nearest_centroids = prefhetch::Client::sort_nearest_centroids(query.data(), nq, cent.data(), nlist, d);
        std::vector<int64_t> probe_ids(nq * nprobe);
        for (uint64_t i = 0; i < nq; i++)
            for (uint64_t p = 0; p < nprobe; p++) probe_ids[i * nprobe + p] = nearest_centroids[i][p].idx;
        // encrypted coarse query
        std::vector<int64_t> qi(query.begin(), query.end());
        std::vector<uint8_t> blob;
        std::vector<uint64_t> offs{0};
        for (uint64_t i = 0; i < nq; i++) {
            std::vector<uint64_t> o;
            const auto b = he.compute_encrypted_coarse_query(qi.data() + i * d, &o);
            for (size_t a = 1; a < o.size(); a++) offs.push_back(blob.size() + o[a]);
            blob.insert(blob.end(), b.begin(), b.end());
        }
        const std::string ereq = prefhetch::Client::coarse_search_encrypted_request(blob, offs, probe_ids.data(), nq, nprobe);
        const std::string eresp = h::coarse_search_encrypted(srv, ereq); // POST /coarsesearch-encrypted
        const auto resp = prefhetch::Client::parse_coarse_search_encrypted_response(eresp);
        std::vector<float> coarse_scores;
        std::vector<int64_t> coarse_idx;
        std::vector<uint64_t> list_sizes;
        int budget = 0;
        he.get_coarse_scores(resp, qi.data(), coarse_scores, coarse_idx, list_sizes, &budget);
        // the plaintext endpoint of the reference returns the same numbers
        std::string preq = "{\"preciseQuery\":";
        h::json::put_matrix(preq, query.data(), nq, d);
        preq += ",\"nearestCentroidIndexes\":";
        h::json::put_matrix(preq, probe_ids.data(), nq, nprobe);
        preq += "}";
        const std::string presp_body = h::coarse_search(srv, preq); // the parsed views point into it
        const auto presp = h::json::object(presp_body);
        if (h::json::vector<float>(h::json::at(presp, "coarseDistanceScores")) != coarse_scores) throw std::runtime_error("decrypted scores differ from the plaintext endpoint");
        if (h::json::vector<int64_t>(h::json::at(presp, "coarseVectorIndexes")) != coarse_idx) throw std::runtime_error("labels differ from the plaintext endpoint");
        // rank, re-rank, compare with brute force over the probed candidates
        const auto coarse = prefhetch::Client::compute_nearest_coarse_vectors(coarse_scores, coarse_idx, list_sizes, coarse_probe);
        std::vector<int64_t> cand(nq * coarse_probe);
        for (uint64_t i = 0; i < nq; i++)
            for (uint64_t j = 0; j < coarse_probe; j++) cand[i * coarse_probe + j] = coarse[i][j].idx;
        std::string sreq = "{\"preciseQuery\":";
        h::json::put_matrix(sreq, query.data(), nq, d);
        sreq += ",\"nearestCoarseVectorIndexes\":";
        h::json::put_matrix(sreq, cand.data(), nq, coarse_probe);
        sreq += "}";
        const std::string sresp_body = h::precise_search(srv, sreq);
        const auto sresp = h::json::object(sresp_body);
        const std::vector<float> pscores = h::json::matrix<float>(h::json::at(sresp, "preciseDistanceScores"), r, c);
        const auto precise = prefhetch::Client::compute_nearest_precise_vectors(pscores.data(), coarse, coarse_probe);
        uint64_t hits = 0;
        for (uint64_t i = 0; i < nq; i++) {
            // exact brute force over every candidate of the probed lists
            std::vector<std::pair<int64_t, int64_t>> bf;
            for (uint64_t p = 0; p < nprobe; p++) {
                const int64_t l = probe_ids[i * nprobe + p];
                for (idx_t o = offsets[l]; o < offsets[l + 1]; o++) {
                    int64_t dist = 0;
                    for (uint32_t k = 0; k < d; k++) {
                        const int64_t df = (int64_t)vectors[(size_t)o * d + k] - (int64_t)query[i * d + k];
                        dist += df * df;
                    }
                    bf.push_back({dist, ids[(size_t)o]});
                }
            }
            std::stable_sort(bf.begin(), bf.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
            for (uint64_t j = 0; j < K; j++) {
                if ((int64_t)precise[i][j].distance != bf[j].first) throw std::runtime_error("re-ranked distance differs from brute force");
                hits += precise[i][j].idx == bf[j].second;
            }
        }
        printf("pf_roundtrip_example ok: %llu queries, %zu candidates decrypted from %zu result ciphertexts, noise budget %d bits, "
               "top-%llu distances exact, %llu of %llu ids in brute-force order\n",
               (unsigned long long)nq, coarse_scores.size(), resp.result_offsets.size() - 1, budget, (unsigned long long)K,
               (unsigned long long)hits, (unsigned long long)(nq * K));
        return 0;
    } catch (const std::exception &ex) {
        fprintf(stderr, "pf_roundtrip_example: %s\n", ex.what());
        return 1;
    }
}
