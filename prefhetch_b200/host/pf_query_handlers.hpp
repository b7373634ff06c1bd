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

#include "pf_server.hpp"

namespace prefhetch::handlers {

// ---- JSON, as far as these endpoints need it ------------------------------------------------------------
namespace json {

inline void skip_ws(std::string_view s, size_t &i) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++;
}

// end (one past) of the value starting at i: string, array, object, or bare token
inline size_t value_end(std::string_view s, size_t i) {
    if (i >= s.size()) throw std::runtime_error("json: unexpected end of input");
    if (s[i] == '"') {
        for (size_t j = i + 1; j < s.size(); j++) {
            if (s[j] == '\\') j++;
            else if (s[j] == '"') return j + 1;
        }
        throw std::runtime_error("json: unterminated string");
    }
    if (s[i] == '[' || s[i] == '{') {
        int depth = 0;
        for (size_t j = i; j < s.size(); j++) {
            if (s[j] == '"') j = value_end(s, j) - 1;
            else if (s[j] == '[' || s[j] == '{') depth++;
            else if (s[j] == ']' || s[j] == '}') {
                if (--depth == 0) return j + 1;
            }
        }
        throw std::runtime_error("json: unbalanced brackets");
    }
    size_t j = i;
    while (j < s.size() && s[j] != ',' && s[j] != ']' && s[j] != '}' && s[j] != ' ' && s[j] != '\n' && s[j] != '\t' && s[j] != '\r') j++;
    if (j == i) throw std::runtime_error("json: value expected");
    return j;
}

// top-level object -> key -> raw text of the value (views into `body`)
inline std::map<std::string, std::string_view> object(std::string_view body) {
    std::map<std::string, std::string_view> out;
    size_t i = 0;
    skip_ws(body, i);
    if (i >= body.size() || body[i] != '{') throw std::runtime_error("json: object expected");
    i++;
    skip_ws(body, i);
    if (i < body.size() && body[i] == '}') return out;
    for (;;) {
        skip_ws(body, i);
        if (i >= body.size() || body[i] != '"') throw std::runtime_error("json: key expected");
        const size_t ke = value_end(body, i);
        std::string key(body.substr(i + 1, ke - i - 2));
        i = ke;
        skip_ws(body, i);
        if (i >= body.size() || body[i] != ':') throw std::runtime_error("json: ':' expected");
        i++;
        skip_ws(body, i);
        const size_t ve = value_end(body, i);
        out.emplace(std::move(key), body.substr(i, ve - i));
        i = ve;
        skip_ws(body, i);
        if (i < body.size() && body[i] == ',') {
            i++;
            continue;
        }
        if (i < body.size() && body[i] == '}') return out;
        throw std::runtime_error("json: ',' or '}' expected");
    }
}

inline std::string_view at(const std::map<std::string, std::string_view> &o, const char *key) {
    auto it = o.find(key);
    if (it == o.end()) throw std::runtime_error(std::string("json: key '") + key + "' not found"); // nlohmann: out_of_range
    return it->second;
}

template <class T> T number(std::string_view tok) {
    T v{};
    const char *b = tok.data(), *e = tok.data() + tok.size();
    if (b < e && *b == '+') b++;
    auto [p, ec] = std::from_chars(b, e, v);
    if (ec != std::errc() || p != e) throw std::runtime_error("json: bad number '" + std::string(tok) + "'");
    return v;
}

// flat array of numbers: [a, b, ...]
template <class T> std::vector<T> vector(std::string_view s) {
    std::vector<T> out;
    size_t i = 0;
    skip_ws(s, i);
    if (i >= s.size() || s[i] != '[') throw std::runtime_error("json: array expected");
    i++;
    skip_ws(s, i);
    if (i < s.size() && s[i] == ']') return out;
    for (;;) {
        skip_ws(s, i);
        const size_t ve = value_end(s, i);
        out.push_back(number<T>(s.substr(i, ve - i)));
        i = ve;
        skip_ws(s, i);
        if (i < s.size() && s[i] == ',') {
            i++;
            continue;
        }
        if (i < s.size() && s[i] == ']') return out;
        throw std::runtime_error("json: ',' or ']' expected");
    }
}

// array of equally long arrays of numbers -> row-major flat vector; rows / cols returned
template <class T> std::vector<T> matrix(std::string_view s, size_t &rows, size_t &cols) {
    std::vector<T> out;
    rows = cols = 0;
    size_t i = 0;
    skip_ws(s, i);
    if (i >= s.size() || s[i] != '[') throw std::runtime_error("json: array of arrays expected");
    i++;
    skip_ws(s, i);
    if (i < s.size() && s[i] == ']') return out;
    for (;;) {
        skip_ws(s, i);
        const size_t ve = value_end(s, i);
        const std::vector<T> row = vector<T>(s.substr(i, ve - i));
        if (rows == 0) cols = row.size();
        else if (row.size() != cols) throw std::runtime_error("json: rows of different lengths"); // std::array would not convert
        out.insert(out.end(), row.begin(), row.end());
        rows++;
        i = ve;
        skip_ws(s, i);
        if (i < s.size() && s[i] == ',') {
            i++;
            continue;
        }
        if (i < s.size() && s[i] == ']') return out;
        throw std::runtime_error("json: ',' or ']' expected");
    }
}

inline std::string string(std::string_view s) {
    size_t i = 0;
    skip_ws(s, i);
    if (i >= s.size() || s[i] != '"') throw std::runtime_error("json: string expected");
    const size_t e = value_end(s, i);
    std::string out;
    for (size_t j = i + 1; j + 1 < e; j++) {
        if (s[j] == '\\' && j + 2 < e) {
            const char c = s[++j];
            out.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c == 'r' ? '\r' : c); // \" \\ \/ and the common escapes
        } else {
            out.push_back(s[j]);
        }
    }
    return out;
}

template <class T> void put_number(std::string &out, T v) {
    char buf[40];
    auto [p, ec] = std::to_chars(buf, buf + sizeof(buf), v); // floats: shortest text that reads back to the same value
    (void)ec;
    out.append(buf, p);
}
template <class T> void put_vector(std::string &out, const T *v, size_t n) {
    out.push_back('[');
    for (size_t i = 0; i < n; i++) {
        if (i) out.push_back(',');
        put_number(out, v[i]);
    }
    out.push_back(']');
}
template <class T> void put_matrix(std::string &out, const T *v, size_t rows, size_t cols) {
    out.push_back('[');
    for (size_t r = 0; r < rows; r++) {
        if (r) out.push_back(',');
        put_vector(out, v + r * cols, cols);
    }
    out.push_back(']');
}

inline std::string base64_encode(const uint8_t *p, size_t n) {
    static const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::string out;
    out.reserve((n + 2) / 3 * 4);
    size_t i = 0;
    for (; i + 2 < n; i += 3) {
        const uint32_t v = (uint32_t)p[i] << 16 | (uint32_t)p[i + 1] << 8 | p[i + 2];
        out.push_back(A[v >> 18]);
        out.push_back(A[(v >> 12) & 63]);
        out.push_back(A[(v >> 6) & 63]);
        out.push_back(A[v & 63]);
    }
    if (i < n) {
        const uint32_t v = (uint32_t)p[i] << 16 | (i + 1 < n ? (uint32_t)p[i + 1] << 8 : 0);
        out.push_back(A[v >> 18]);
        out.push_back(A[(v >> 12) & 63]);
        out.push_back(i + 1 < n ? A[(v >> 6) & 63] : '=');
        out.push_back('=');
    }
    return out;
}

inline std::vector<uint8_t> base64_decode(std::string_view s) {
    static const auto table = [] {
        std::vector<int8_t> t(256, -1);
        const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) t[(uint8_t)A[i]] = (int8_t)i;
        return t;
    }();
    std::vector<uint8_t> out;
    out.reserve(s.size() / 4 * 3);
    uint32_t acc = 0;
    int bits = 0;
    for (char c : s) {
        if (c == '=' || c == '\n' || c == '\r') continue;
        const int8_t v = table[(uint8_t)c];
        if (v < 0) throw std::runtime_error("base64: invalid character");
        acc = acc << 6 | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
        }
    }
    return out;
}

} // namespace json

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
