// pf_query_handlers.hpp — the bodies of the reference's HTTP handlers over the GPU engine.
//
// The reference's controller (ref: src/server/controllers/Query.cc:10-98) does three things per endpoint: parse the
// JSON body into arrays (nlohmann::json), call the `Server` singleton, dump the result as JSON.  The functions below
// are those bodies with `prefhetch::Server` (pf_server.hpp) in place of the FAISS-backed class: request body in,
// response body out, same field names, run-time shapes instead of std::array<..., NQUERY>.  A Drogon handler
// becomes three lines:
//     resp->setBody(prefhetch::handlers::coarse_search(gpu, req->body()));
// nlohmann-json and Drogon are not in this image, so the envelope is read and written by a small
// schema-specific JSON codec (objects whose values are numbers, strings or nested arrays of numbers — all these
// endpoints exchange); where nlohmann is available it can replace the codec without touching the handler bodies.
// Errors are std::runtime_error, as in the reference (uncaught in its handlers -> HTTP 500).
//
//   GET  /query               -> [[f32 x d] x nlist]                                   (Query.cc:10-27)
//   POST /coarsesearch        {preciseQuery:[[..]], nearestCentroidIndexes:[[..]]}
//                             -> {coarseDistanceScores, coarseVectorIndexes, listSizesPerQuery}   (Query.cc:29-63)
//   POST /precisesearch       {preciseQuery, nearestCoarseVectorIndexes} -> {preciseDistanceScores:[[..]]}  (Query.cc:65-98)
//   POST /galoiskeys          (additive)  {galoisKeys: base64 of the client's SEAL GaloisKeys stream} -> {galoisKeysBytes}
//   POST /coarsesearch-encrypted (additive)  {queryCiphertexts: base64 of the SEAL streams back to back,
//                             ctOffsets:[..], nearestCentroidIndexes:[[..]]}
//                             -> {resultCiphertexts: base64, resultOffsets, resultsPerQuery, coarseVectorIndexes,
//                                 listSizesPerQuery, probedListSizes, resultBytes}
//   POST /precise-vector-pir  is a plain gather of raw vectors (Query.cc:100-127): out of scope, stays as it is.
#pragma once
#include <charconv>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "pf_json.hpp"
#include "pf_server.hpp"

namespace prefhetch::handlers {

// ---- handler bodies ---------------------------------------------------------------------------------------

// GET /query (ref: Query.cc:10-27): all centroids as [[f32 x d] x nlist]
inline std::string query(const Server &srv) {
    std::vector<float> cent;
    srv.retrieve_centroids(cent);
    const size_t d = srv.dim();
    std::string out;
    json::put_matrix(out, cent.data(), cent.size() / d, d);
    return out;
}

// POST /coarsesearch (ref: Query.cc:29-63 -> Server::coarseSearch, src/server/server_lib.cpp:111-138)
inline std::string coarse_search(const Server &srv, std::string_view body) {
    const auto req = json::object(body);
    size_t nq = 0, d = 0, nq2 = 0, nprobe = 0;
    const std::vector<float> precise_query = json::matrix<float>(json::at(req, "preciseQuery"), nq, d);
    const std::vector<idx_t> nearest = json::matrix<idx_t>(json::at(req, "nearestCentroidIndexes"), nq2, nprobe);
    if (d != srv.dim() || nq != nq2 || !nprobe) throw std::runtime_error("coarsesearch: preciseQuery / nearestCentroidIndexes shapes do not match");
    std::vector<float> scores;
    std::vector<idx_t> indexes;
    std::vector<size_t> sizes;
    srv.coarseSearch(precise_query, nearest, (uint32_t)nprobe, scores, indexes, sizes);
    std::string out = "{\"coarseDistanceScores\":";
    json::put_vector(out, scores.data(), scores.size());
    out += ",\"coarseVectorIndexes\":";
    json::put_vector(out, indexes.data(), indexes.size());
    out += ",\"listSizesPerQuery\":";
    json::put_vector(out, sizes.data(), sizes.size());
    out.push_back('}');
    return out;
}

// POST /precisesearch (ref: Query.cc:65-98 -> Server::preciseSearch, src/server/server_lib.cpp:140-167)
inline std::string precise_search(const Server &srv, std::string_view body) {
    const auto req = json::object(body);
    size_t nq = 0, d = 0, nq2 = 0, probe = 0;
    const std::vector<float> precise_query = json::matrix<float>(json::at(req, "preciseQuery"), nq, d);
    const std::vector<idx_t> ids = json::matrix<idx_t>(json::at(req, "nearestCoarseVectorIndexes"), nq2, probe);
    if (d != srv.dim() || nq != nq2 || !probe) throw std::runtime_error("precisesearch: preciseQuery / nearestCoarseVectorIndexes shapes do not match");
    std::vector<float> scores;
    srv.preciseSearch(precise_query, ids, (uint32_t)probe, scores);
    std::string out = "{\"preciseDistanceScores\":";
    json::put_matrix(out, scores.data(), nq, probe);
    out.push_back('}');
    return out;
}

// POST /galoiskeys (additive): the client's SEAL GaloisKeys stream (full or seeded, any compr_mode the engine reads),
// base64 in {"galoisKeys": "..."} — once per client session, before the first encrypted search.  The engine holds one
// key set at a time (one client); multi-tenant key management is the server application's business.
inline std::string galois_keys(Server &srv, std::string_view body) {
    const auto req = json::object(body);
    const std::vector<uint8_t> blob = json::base64_decode(json::string(json::at(req, "galoisKeys")));
    srv.loadGaloisKeys(blob);
    std::string out = "{\"galoisKeysBytes\":";
    json::put_number(out, (uint64_t)blob.size());
    out.push_back('}');
    return out;
}

// POST /coarsesearch-encrypted (additive; the TODO of include/client/client_lib.h:14,28-30 and
// src/server/controllers/Query.h:17-18): SEAL-serialized query ciphertexts in, result ciphertexts out
inline std::string coarse_search_encrypted(const Server &srv, std::string_view body) {
    const auto req = json::object(body);
    const std::vector<uint8_t> blob = json::base64_decode(json::string(json::at(req, "queryCiphertexts")));
    const std::vector<uint64_t> offsets = json::vector<uint64_t>(json::at(req, "ctOffsets"));
    size_t nq = 0, nprobe = 0;
    const std::vector<idx_t> nearest = json::matrix<idx_t>(json::at(req, "nearestCentroidIndexes"), nq, nprobe);
    if (!nq || !nprobe || offsets.size() < 2 || (offsets.size() - 1) % nq) throw std::runtime_error("coarsesearch-encrypted: ctOffsets does not describe a whole number of ciphertexts per query");
    EncryptedCoarseResult r;
    srv.coarseSearchEncrypted(nq, blob, offsets, nearest, (uint32_t)nprobe, r);
    std::string out = "{\"resultCiphertexts\":\"";
    out += json::base64_encode(r.ciphertexts.data(), r.ciphertexts.size());
    out += "\",\"resultOffsets\":";
    json::put_vector(out, r.result_offsets.data(), r.result_offsets.size());
    out += ",\"resultBytes\":";
    json::put_number(out, (uint64_t)srv.resultSerializedSize());
    out += ",\"resultsPerQuery\":";
    json::put_vector(out, r.results_per_query.data(), r.results_per_query.size());
    out += ",\"coarseVectorIndexes\":";
    json::put_vector(out, r.coarse_vector_indexes.data(), r.coarse_vector_indexes.size());
    out += ",\"listSizesPerQuery\":";
    json::put_vector(out, r.list_sizes_per_query.data(), r.list_sizes_per_query.size());
    out += ",\"probedListSizes\":";
    json::put_matrix(out, r.probed_sizes.data(), nq, nprobe);
    out.push_back('}');
    return out;
}

} // namespace prefhetch::handlers
