// ref_shim_client.cpp — C entry points onto the REFERENCE's own client functions, compiled from
// /root/reference/src/client/client_lib.cpp where it lies (oracle/ref_build/Makefile).  Test infrastructure: the
// oracle's restatements of these functions are pinned against what this library returns (tests/golden/).
// Shapes are the reference's compile-time constants (include/common/client_server_utils.h:10-20).
#include <array>
#include <chrono>
#include <string>
#include <vector>

#include "client_lib.h"

static std::vector<std::string> g_log;
void pf_ref_log_line(const std::string &line) { g_log.push_back(line); }

extern "C" {

void ref_constants(int64_t *out /*[8]*/) {
    const int64_t c[8] = {PRECISE_VECTOR_DIMENSIONS, NPROBE, COARSE_PROBE, K, NBASE, NQUERY, NLIST, SUB_QUANTIZERS};
    for (int i = 0; i < 8; i++) out[i] = c[i];
}

// ref: src/client/client_lib.cpp:49-81.  idx_out / dist_out [NQUERY][nlist]: every centroid, in the order the
// reference's sort leaves them
void ref_sort_nearest_centroids(const float *query /*[NQUERY][128]*/, const float *centroids /*[nlist][128]*/, int64_t nlist, int64_t *idx_out,
                                float *dist_out) {
    std::array<std::array<float, PRECISE_VECTOR_DIMENSIONS>, NQUERY> q;
    for (int i = 0; i < NQUERY; i++)
        for (int k = 0; k < PRECISE_VECTOR_DIMENSIONS; k++) q[i][k] = query[i * PRECISE_VECTOR_DIMENSIONS + k];
    std::vector<std::array<float, PRECISE_VECTOR_DIMENSIONS>> cent(nlist);
    for (int64_t j = 0; j < nlist; j++)
        for (int k = 0; k < PRECISE_VECTOR_DIMENSIONS; k++) cent[j][k] = centroids[j * PRECISE_VECTOR_DIMENSIONS + k];
    std::array<std::vector<DistanceIndexData>, NQUERY> nearest;
    sort_nearest_centroids(q, cent, nearest);
    for (int i = 0; i < NQUERY; i++)
        for (int64_t j = 0; j < nlist; j++) {
            idx_out[i * nlist + j] = nearest[i][j].idx;
            dist_out[i * nlist + j] = nearest[i][j].distance;
        }
}

// the same call `reps` times on the same inputs; returns seconds per call (bench.py --impl reference, configs[0])
double ref_sort_nearest_centroids_timed(const float *query, const float *centroids, int64_t nlist, int reps) {
    std::array<std::array<float, PRECISE_VECTOR_DIMENSIONS>, NQUERY> q;
    for (int i = 0; i < NQUERY; i++)
        for (int k = 0; k < PRECISE_VECTOR_DIMENSIONS; k++) q[i][k] = query[i * PRECISE_VECTOR_DIMENSIONS + k];
    std::vector<std::array<float, PRECISE_VECTOR_DIMENSIONS>> cent(nlist);
    for (int64_t j = 0; j < nlist; j++)
        for (int k = 0; k < PRECISE_VECTOR_DIMENSIONS; k++) cent[j][k] = centroids[j * PRECISE_VECTOR_DIMENSIONS + k];
    const auto t0 = std::chrono::steady_clock::now();
    volatile float sink = 0;
    for (int r = 0; r < reps; r++) {
        std::array<std::vector<DistanceIndexData>, NQUERY> nearest;
        sort_nearest_centroids(q, cent, nearest);
        sink = sink + nearest[0][0].distance;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (reps > 0 ? reps : 1);
}

// ref: src/client/client_lib.cpp:122-156.  Returns 0, or 1 when the reference throws (fewer than COARSE_PROBE
// candidates for a query).  idx_out / dist_out packed per query like the input, each query's range sorted.
int ref_compute_nearest_coarse_vectors(const float *scores, const int64_t *indexes, const uint64_t *list_sizes /*[NQUERY]*/, int64_t *idx_out,
                                       float *dist_out) {
    size_t total = 0;
    std::array<size_t, NQUERY> sizes;
    for (int i = 0; i < NQUERY; i++) total += (sizes[i] = (size_t)list_sizes[i]);
    std::vector<float> s(scores, scores + total);
    std::vector<faiss_idx_t> ix(indexes, indexes + total);
    std::array<std::vector<DistanceIndexData>, NQUERY> nearest;
    try {
        compute_nearest_coarse_vectors(s, ix, sizes, nearest);
    } catch (const std::runtime_error &) {
        return 1;
    }
    size_t o = 0;
    for (int i = 0; i < NQUERY; i++)
        for (const DistanceIndexData &e : nearest[i]) {
            idx_out[o] = e.idx;
            dist_out[o++] = e.distance;
        }
    return 0;
}

// ref: src/client/client_lib.cpp:189-209.  coarse_ids [NQUERY][COARSE_PROBE] = the ids the precise scores belong to
void ref_compute_nearest_precise_vectors(const float *precise_scores /*[NQUERY][COARSE_PROBE]*/, const int64_t *coarse_ids, int64_t *idx_out,
                                         float *dist_out) {
    std::array<std::array<float, COARSE_PROBE>, NQUERY> ps;
    std::array<std::vector<DistanceIndexData>, NQUERY> coarse;
    for (int i = 0; i < NQUERY; i++)
        for (int j = 0; j < COARSE_PROBE; j++) {
            ps[i][j] = precise_scores[i * COARSE_PROBE + j];
            coarse[i].push_back(DistanceIndexData{0.0f, coarse_ids[i * COARSE_PROBE + j]});
        }
    std::array<std::array<DistanceIndexData, COARSE_PROBE>, NQUERY> out;
    compute_nearest_precise_vectors(ps, coarse, out);
    for (int i = 0; i < NQUERY; i++)
        for (int j = 0; j < COARSE_PROBE; j++) {
            idx_out[i * COARSE_PROBE + j] = out[i][j].idx;
            dist_out[i * COARSE_PROBE + j] = out[i][j].distance;
        }
}

// ref: src/client/client_lib.cpp:243-337 benchmark_results: reads ../sift/siftsmall/siftsmall_groundtruth.ivecs
// relative to the working directory and LOGS its numbers; the log lines (format string | arg | arg ...) are
// returned, newline separated.  Returns the number of bytes needed.
size_t ref_benchmark_results(const int64_t *observed /*[NQUERY][K]*/, char *out, size_t cap) {
    std::array<std::array<faiss_idx_t, K>, NQUERY> obs;
    for (int i = 0; i < NQUERY; i++)
        for (int j = 0; j < K; j++) obs[i][j] = observed[i * K + j];
    g_log.clear();
    benchmark_results(obs);
    std::string all;
    for (const std::string &l : g_log) all += l + "\n";
    if (out && cap) {
        const size_t n = all.size() < cap - 1 ? all.size() : cap - 1;
        all.copy(out, n);
        out[n] = 0;
    }
    return all.size() + 1;
}

} // extern "C"
