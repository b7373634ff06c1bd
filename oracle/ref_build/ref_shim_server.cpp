// ref_shim_server.cpp — C entry points onto the REFERENCE's own Server::init_index / Server::preciseSearch, compiled
// from /root/reference/src/server/server_lib.cpp where it lies, against compile-only stand-ins for FAISS / Drogon
// (oracle/ref_build/stubs: train / add / write_index do nothing).  Test infrastructure, see ref_shim_client.cpp.
#include <array>
#include <chrono>
#include <string>
#include <vector>

#include "server_lib.h"

static std::vector<std::string> g_log;
void pf_ref_log_line(const std::string &line) { g_log.push_back(line); }

extern "C" {

// Server::init_index (ref: src/server/server_lib.cpp:55-99) reads ../sift/siftsmall/siftsmall_{learn,base}.fvecs
// relative to the working directory; Server::preciseSearch (ref: :140-167) on NQUERY queries x COARSE_PROBE ids.
// Returns 0, or 1 when the reference throws.
int ref_precise_search(const float *query /*[NQUERY][128]*/, const int64_t *ids /*[NQUERY][COARSE_PROBE]*/, float *out /*[NQUERY][COARSE_PROBE]*/) {
    try {
        Server srv;
        srv.init_index();
        std::array<std::array<float, PRECISE_VECTOR_DIMENSIONS>, NQUERY> q;
        std::array<std::array<faiss::idx_t, COARSE_PROBE>, NQUERY> ix;
        std::array<std::array<float, COARSE_PROBE>, NQUERY> scores;
        for (int i = 0; i < NQUERY; i++) {
            for (int k = 0; k < PRECISE_VECTOR_DIMENSIONS; k++) q[i][k] = query[i * PRECISE_VECTOR_DIMENSIONS + k];
            for (int j = 0; j < COARSE_PROBE; j++) ix[i][j] = ids[i * COARSE_PROBE + j];
        }
        srv.preciseSearch(q, ix, scores);
        for (int i = 0; i < NQUERY; i++)
            for (int j = 0; j < COARSE_PROBE; j++) out[i * COARSE_PROBE + j] = scores[i][j];
        return 0;
    } catch (const std::exception &) {
        return 1;
    }
}

// Server::preciseSearch `reps` times on one Server (init_index once, outside the timed loop); seconds per call, < 0 on
// a throw (bench.py --impl reference, configs[0])
double ref_precise_search_timed(const float *query, const int64_t *ids, int reps) {
    try {
        Server srv;
        srv.init_index();
        std::array<std::array<float, PRECISE_VECTOR_DIMENSIONS>, NQUERY> q;
        std::array<std::array<faiss::idx_t, COARSE_PROBE>, NQUERY> ix;
        std::array<std::array<float, COARSE_PROBE>, NQUERY> scores;
        for (int i = 0; i < NQUERY; i++) {
            for (int k = 0; k < PRECISE_VECTOR_DIMENSIONS; k++) q[i][k] = query[i * PRECISE_VECTOR_DIMENSIONS + k];
            for (int j = 0; j < COARSE_PROBE; j++) ix[i][j] = ids[i * COARSE_PROBE + j];
        }
        const auto t0 = std::chrono::steady_clock::now();
        volatile float sink = 0;
        for (int r = 0; r < reps; r++) {
            srv.preciseSearch(q, ix, scores);
            sink = sink + scores[0][0];
        }
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (reps > 0 ? reps : 1);
    } catch (const std::exception &) {
        return -1.0;
    }
}

} // extern "C"
