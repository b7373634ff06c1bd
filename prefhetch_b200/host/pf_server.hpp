// pf_server.hpp — C++ host mirror of the reference `Server` class over the C ABI
// (include/prefhetch_b200.h).  Same method names and argument meaning as the reference
// (ref: include/server/server_lib.h:25-49, src/server/server_lib.cpp:101-167) with run-time
// shapes (std::vector / std::span) instead of compile-time std::array, and std::runtime_error
// where the reference throws it (ref: src/server/server_lib.cpp:66,94).  Header-only; link with
// -lprefhetch_b200.  This is the code a maintainer drops into src/server/server_lib.cpp
// (see INTEGRATION.md); it needs neither FAISS nor SEAL on the server.
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/prefhetch_b200.h"
#include "pf_faiss_io.hpp"

namespace prefhetch {

using idx_t = int64_t; // ref: faiss_idx_t, include/common/client_server_utils.h:22

struct EncryptedCoarseResult {
    std::vector<uint8_t> ciphertexts;        // SEAL-serialized result ciphertexts in aligned slots (pf_result_slot_size)
    std::vector<uint64_t> result_offsets;    // [nresults+1]
    std::vector<uint64_t> results_per_query; // [nq]
    std::vector<idx_t> coarse_vector_indexes; // ids of the probed lists, packed per query (ref field name)
    std::vector<uint64_t> list_sizes_per_query;
    std::vector<uint64_t> probed_sizes;      // [nq][nprobe]
    uint64_t nresults = 0, out_bytes = 0;
};

class Server {
  public:
    // ref: Server::Server (src/server/server_lib.cpp:32-46); parameters are run-time here
    Server(uint32_t dim, uint64_t poly_degree, const std::vector<uint64_t> &primes, uint64_t plain_modulus,
           uint32_t query_cts = 1, uint32_t partial_g = 8, int device = 0, uint32_t rank = 0, uint32_t world = 1,
           uint32_t result_limbs = 0)
        : m_Dim(dim) {
        pf_params p{};
        p.struct_size = sizeof(pf_params);
        p.device = device;
        p.poly_degree = poly_degree;
        p.num_primes = static_cast<uint32_t>(primes.size());
        p.dim = dim;
        for (size_t i = 0; i < primes.size() && i < PF_MAX_PRIMES; i++) p.primes[i] = primes[i];
        p.plain_modulus = plain_modulus;
        p.query_cts = query_cts;
        p.partial_g = partial_g;
        p.rank = rank;
        p.world = world;
        p.result_limbs = result_limbs; // 0 = full level; 1 = what bench.py ships (SEAL mod_switch_to before save)
        pf_engine *e = nullptr;
        if (pf_engine_create(&p, &e) != PF_OK) throw std::runtime_error(pf_last_error(nullptr));
        m_Engine.reset(e);
    }

    // ref: Server::init_index hands the trained IVF index to the searcher (src/server/server_lib.cpp:55-99)
    void init_index(uint64_t nlist, const float *centroids, const idx_t *list_offsets, const idx_t *ids,
                    const float *vectors) {
        m_Nlist = nlist;
        m_ListOffsets.assign(list_offsets, list_offsets + nlist + 1);
        check(pf_load_index(m_Engine.get(), nlist, centroids, list_offsets, ids, vectors));
        check(pf_get_index_info(m_Engine.get(), &m_Info));
    }

    // ref: the cached-file branch of Server::init_index (src/server/server_lib.cpp:88-99): centroids and
    // inverted lists from the .faiss file, raw vectors from the base set addressed by id (:154-156)
    void init_index(const std::string &faiss_path, std::span<const float> dataset_base) {
        const IvfFile f = read_ivfpq_file(faiss_path);
        if (f.d != m_Dim) throw std::runtime_error("index dimension does not match PRECISE_VECTOR_DIMENSIONS");
        std::vector<float> vectors(f.ntotal * m_Dim);
        for (uint64_t i = 0; i < f.ntotal; i++) {
            const uint64_t row = static_cast<uint64_t>(f.ids[i]);
            if (f.ids[i] < 0 || (row + 1) * m_Dim > dataset_base.size()) throw std::runtime_error("id outside the base set");
            std::copy_n(dataset_base.data() + row * m_Dim, m_Dim, vectors.data() + i * m_Dim);
        }
        init_index(f.nlist, f.centroids.data(), f.list_offsets.data(), f.ids.data(), vectors.data());
        // the file's product quantizer (8-bit sub-quantizers, one code byte each: what the reference builds) serves coarseSearchPQ
        if (f.pq_nbits == 8 && f.pq_M && f.code_size == f.pq_M && f.pq_centroids.size() == size_t(256) * m_Dim)
            loadProductQuantizer(static_cast<uint32_t>(f.pq_M), 8, f.pq_centroids, f.codes);
    }

    // ref: Server::retrieve_centroids (src/server/server_lib.cpp:101-109)
    void retrieve_centroids(std::vector<float> &centroids) const {
        centroids.resize(m_Nlist * m_Dim);
        check(pf_retrieve_centroids(m_Engine.get(), centroids.data(), centroids.size()));
    }

    // stage 1 on the server (the reference runs it on the client: src/client/client_lib.cpp:50-81)
    void coarseQuantize(std::span<const float> precise_query, uint32_t nprobe, std::vector<idx_t> &nearest) const {
        const uint64_t nq = precise_query.size() / m_Dim;
        nearest.resize(nq * nprobe);
        check(pf_coarse_quantize(m_Engine.get(), nq, precise_query.data(), nprobe, nearest.data(), nullptr));
    }

    // ref: Server::coarseSearch (src/server/server_lib.cpp:111-138) — same outputs, sized exactly
    void coarseSearch(std::span<const float> precise_query, std::span<const idx_t> nearest_centroid_idx,
                      uint32_t nprobe, std::vector<float> &coarse_distance_scores,
                      std::vector<idx_t> &coarse_distance_indexes, std::vector<size_t> &list_sizes_per_query) const {
        const uint64_t nq = precise_query.size() / m_Dim;
        std::vector<uint64_t> sizes(nq);
        uint64_t total = 0;
        int rc = pf_search_lists_plain(m_Engine.get(), nq, precise_query.data(), nearest_centroid_idx.data(), nprobe,
                                       nullptr, nullptr, 0, sizes.data(), &total);
        if (rc != PF_OK && rc != PF_ERR_CAPACITY) check(rc);
        coarse_distance_scores.resize(total);
        coarse_distance_indexes.resize(total);
        check(pf_search_lists_plain(m_Engine.get(), nq, precise_query.data(), nearest_centroid_idx.data(), nprobe,
                                    coarse_distance_scores.data(), coarse_distance_indexes.data(), total, sizes.data(),
                                    &total));
        list_sizes_per_query.assign(sizes.begin(), sizes.end());
    }

    // The product quantizer of the loaded index (the IndexIVFPQ of ref: src/server/server_lib.cpp:34-36; from the
    // .faiss file: host/pf_faiss_io.hpp) and Server::coarseSearch with the distance the reference's FAISS fork
    // computes TODAY — PQ-ADC over every code of the given lists (the search_encrypted call at
    // ref: src/server/server_lib.cpp:126-130); outputs as coarseSearch
    void loadProductQuantizer(uint32_t sub_quantizers, uint32_t sub_quantizer_bits, std::span<const float> pq_centroids,
                              std::span<const uint8_t> codes) {
        check(pf_load_pq(m_Engine.get(), sub_quantizers, sub_quantizer_bits, pq_centroids.data(), codes.data()));
    }
    void coarseSearchPQ(std::span<const float> precise_query, std::span<const idx_t> nearest_centroid_idx, uint32_t nprobe,
                        std::vector<float> &coarse_distance_scores, std::vector<idx_t> &coarse_distance_indexes,
                        std::vector<size_t> &list_sizes_per_query) const {
        const uint64_t nq = precise_query.size() / m_Dim;
        std::vector<uint64_t> sizes(nq);
        uint64_t total = 0;
        int rc = pf_search_lists_pq(m_Engine.get(), nq, precise_query.data(), nearest_centroid_idx.data(), nprobe, nullptr,
                                    nullptr, 0, sizes.data(), &total);
        if (rc != PF_OK && rc != PF_ERR_CAPACITY) check(rc);
        coarse_distance_scores.resize(total);
        coarse_distance_indexes.resize(total);
        check(pf_search_lists_pq(m_Engine.get(), nq, precise_query.data(), nearest_centroid_idx.data(), nprobe,
                                 coarse_distance_scores.data(), coarse_distance_indexes.data(), total, sizes.data(), &total));
        list_sizes_per_query.assign(sizes.begin(), sizes.end());
    }

    // ref: Server::preciseSearch (src/server/server_lib.cpp:140-167)
    void preciseSearch(std::span<const float> precise_query, std::span<const idx_t> nearest_coarse_vector_idx,
                       uint32_t coarse_probe, std::vector<float> &precise_distance_scores) const {
        const uint64_t nq = precise_query.size() / m_Dim;
        precise_distance_scores.resize(nq * coarse_probe);
        check(pf_precise_search(m_Engine.get(), nq, precise_query.data(), nearest_coarse_vector_idx.data(),
                                coarse_probe, precise_distance_scores.data()));
    }

    // SEAL-serialized GaloisKeys of the client: full or seeded (Serializable<GaloisKeys>), compr_mode none / zlib / zstd
    void loadGaloisKeys(std::span<const uint8_t> blob) { check(pf_load_galois_keys(m_Engine.get(), blob.data(), blob.size())); }

    // one Galois key from raw words [L][2][k][N] (GaloisKeys::key(galois_elt) data, NTT form)
    void setGaloisKey(uint32_t galois_elt, std::span<const uint64_t> words) {
        check(pf_set_galois_key(m_Engine.get(), galois_elt, words.data()));
    }
    uint32_t galoisEltFromStep(int step) const { return pf_galois_elt_from_step(m_Engine.get(), step); }

    // encrypted variant of coarseSearch: SEAL-serialized query ciphertexts in, result ciphertexts out
    void coarseSearchEncrypted(uint64_t nq, std::span<const uint8_t> query_cts, std::span<const uint64_t> ct_offsets,
                               std::span<const idx_t> nearest_centroid_idx, uint32_t nprobe,
                               EncryptedCoarseResult &out) const {
        collect(submitSearchEncrypted(nq, query_cts, ct_offsets, nearest_centroid_idx, nprobe, out), out);
    }

    // the same in two halves (pf_search_submit / pf_search_collect): a handler thread submits request i+1
    // before it collects request i, so the GPU never waits for a download.  `out`, `query_cts` stay alive
    // and untouched until collect() returns.
    uint64_t submitSearchEncrypted(uint64_t nq, std::span<const uint8_t> query_cts, std::span<const uint64_t> ct_offsets,
                                   std::span<const idx_t> nearest_centroid_idx, uint32_t nprobe,
                                   EncryptedCoarseResult &out) const {
        uint64_t max_results = 0, max_labels = 0;
        for (idx_t l : nearest_centroid_idx) {
            if (l < 0 || static_cast<uint64_t>(l) >= m_Nlist) throw std::runtime_error("list id out of range");
            const uint64_t n = static_cast<uint64_t>(m_ListOffsets[l + 1] - m_ListOffsets[l]);
            max_results += (n + m_Info.C - 1) / m_Info.C;
            max_labels += n;
        }
        const size_t ct_bytes = pf_result_slot_size(m_Engine.get());
        out.ciphertexts.resize(max_results * ct_bytes);
        out.result_offsets.resize(max_results + 1);
        out.results_per_query.resize(nq);
        out.coarse_vector_indexes.resize(max_labels);
        out.list_sizes_per_query.resize(nq);
        out.probed_sizes.resize(nq * nprobe);
        pf_search_stats st{};
        uint64_t ticket = 0;
        check(pf_search_submit(m_Engine.get(), nq, query_cts.data(), query_cts.size(), ct_offsets.data(),
                               nearest_centroid_idx.data(), nprobe, out.ciphertexts.data(), out.ciphertexts.size(),
                               out.result_offsets.data(), max_results, out.results_per_query.data(),
                               out.coarse_vector_indexes.data(), max_labels, out.list_sizes_per_query.data(),
                               out.probed_sizes.data(), &st, &ticket));
        out.nresults = st.nresults;
        out.out_bytes = st.out_bytes;
        return ticket;
    }

    void collect(uint64_t ticket, EncryptedCoarseResult &out) const {
        check(pf_search_collect(m_Engine.get(), ticket));
        out.ciphertexts.resize(out.out_bytes);
        out.result_offsets.resize(out.nresults + 1);
        uint64_t labels = 0;
        for (uint64_t s : out.list_sizes_per_query) labels += s;
        out.coarse_vector_indexes.resize(labels);
    }

    size_t resultSerializedSize() const { return pf_result_serialized_size(m_Engine.get()); }

    uint32_t dim() const { return m_Dim; }
    uint64_t nlist() const { return m_Nlist; }
    pf_engine *handle() const { return m_Engine.get(); }
    const pf_index_info &info() const { return m_Info; }

  private:
    void check(int rc) const {
        if (rc != PF_OK) throw std::runtime_error(pf_last_error(m_Engine.get()));
    }
    struct Deleter {
        void operator()(pf_engine *e) const { pf_engine_destroy(e); }
    };
    std::unique_ptr<pf_engine, Deleter> m_Engine;
    uint32_t m_Dim;
    uint64_t m_Nlist = 0;
    std::vector<idx_t> m_ListOffsets;
    pf_index_info m_Info{};
};

} // namespace prefhetch
