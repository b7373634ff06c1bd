// pf_handlers_check.cpp — CPU check of the JSON envelope of the handler bodies (pf_query_handlers.hpp): the codec
// reads the reference's request shapes (ref: src/server/controllers/Query.cc:34-42, :71-83) and what it writes reads
// back to the same numbers.  No engine call (no GPU here): the handler bodies themselves run on the GPU box through
// `pf_server_check --handlers`.  Built by __graft_entry__.build(); run by tests/test_abi.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "pf_query_handlers.hpp"

namespace js = prefhetch::handlers::json;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::fprintf(stderr, "check failed at line %d: %s\n", __LINE__, #cond); \
            return 1;                                                      \
        }                                                                  \
    } while (0)

template <class F> static bool throws(F f) {
    try {
        f();
    } catch (const std::runtime_error &) {
        return true;
    }
    return false;
}

int main() {
    // the reference's request: NQUERY = 5 queries of 128 floats, NPROBE = 20 ids each
    std::mt19937 rng(7);
    const size_t nq = 5, d = 128, nprobe = 20;
    std::vector<float> q(nq * d);
    for (auto &x : q) x = (float)(rng() % 256);
    q[3] = 0.1f;
    q[4] = -1.5e-7f;
    q[5] = 16777216.0f;
    std::vector<prefhetch::idx_t> idx(nq * nprobe);
    for (auto &x : idx) x = (prefhetch::idx_t)(rng() % 256);
    std::string body = "{ \"preciseQuery\" : ";
    js::put_matrix(body, q.data(), nq, d);
    body += ",\n \"nearestCentroidIndexes\":";
    js::put_matrix(body, idx.data(), nq, nprobe);
    body += " }";
    const auto req = js::object(body);
    CHECK(req.size() == 2);
    size_t r = 0, c = 0;
    const auto q2 = js::matrix<float>(js::at(req, "preciseQuery"), r, c);
    CHECK(r == nq && c == d && q2 == q); // shortest round-trip text: bit-identical floats
    const auto i2 = js::matrix<prefhetch::idx_t>(js::at(req, "nearestCentroidIndexes"), r, c);
    CHECK(r == nq && c == nprobe && i2 == idx);
    // what nlohmann would print for the same numbers also reads back (floats as doubles' shortest text, exponents)
    const auto v = js::vector<float>("[0.10000000149011612, 1e3, -2.5E-1, 7, 255.0]");
    CHECK(v.size() == 5 && v[0] == 0.1f && v[1] == 1000.0f && v[2] == -0.25f && v[3] == 7.0f && v[4] == 255.0f);
    CHECK(js::vector<int64_t>("[ ]").empty());
    CHECK(js::vector<uint64_t>("[18446744073709551615]")[0] == 18446744073709551615ull);
    // a response like the reference's reads back too
    std::string resp = "{\"coarseDistanceScores\":";
    js::put_vector(resp, q.data(), 10);
    resp += ",\"listSizesPerQuery\":[3,4,5,0,1]}";
    const auto ro = js::object(resp);
    CHECK(js::vector<float>(js::at(ro, "coarseDistanceScores")).size() == 10);
    CHECK(js::vector<size_t>(js::at(ro, "listSizesPerQuery"))[2] == 5);
    // base64, every tail length
    for (size_t n = 0; n < 70; n++) {
        std::vector<uint8_t> raw(n);
        for (auto &b : raw) b = (uint8_t)rng();
        const std::string enc = js::base64_encode(raw.data(), raw.size());
        CHECK(enc.size() == (n + 2) / 3 * 4 && js::base64_decode(enc) == raw);
    }
    CHECK(js::base64_encode(reinterpret_cast<const uint8_t *>("Man"), 3) == "TWFu");
    CHECK(js::base64_encode(reinterpret_cast<const uint8_t *>("Ma"), 2) == "TWE=");
    CHECK(js::string(" \"a\\\"b\\\\c\" ") == "a\"b\\c");
    std::string so = "{\"queryCiphertexts\":\"QUJD\",\"ctOffsets\":[0,3]}";
    CHECK(js::base64_decode(js::string(js::at(js::object(so), "queryCiphertexts"))) == std::vector<uint8_t>({'A', 'B', 'C'}));
    // malformed input: exceptions (std::runtime_error, as the reference's handlers let nlohmann's propagate), no crash
    CHECK(throws([] { js::object("[1,2]"); }));
    CHECK(throws([] { js::object("{\"a\":[1,2}"); }));
    CHECK(throws([] { js::object("{\"a\" 1}"); }));
    CHECK(throws([&] { js::at(req, "nearestCoarseVectorIndexes"); }));
    CHECK(throws([] { size_t a, b; js::matrix<float>("[[1,2],[3]]", a, b); }));   // ragged: std::array would not convert
    CHECK(throws([] { js::vector<int64_t>("[1,x]"); }));
    CHECK(throws([] { js::vector<int64_t>("[1.5]"); }));
    CHECK(throws([] { js::vector<float>("[1,2"); }));
    CHECK(throws([] { js::base64_decode("ab*d"); }));
    CHECK(throws([] { js::string("\"abc"); }));
    std::printf("pf_handlers_check ok\n");
    return 0;
}
