// pf_client_check.cpp — drives prefhetch::Client (pf_client.hpp) on files, for tests/test_client.py (CPU: the
// oracle plays the server) and tests/test_gpu_parity.py (the engine is the server).  No CUDA, no SEAL.
//   pf_client_check keygen  <dir>   params.txt, seed.bin, queries.i64 -> sk.i8, galois_keys.bin (seeded), galois_keys_full.bin,
//                                   queries_seeded.bin/.off, queries_full.bin/.off, encode_probe.u64
//   pf_client_check nearest <dir>   params.txt, queries.f32, centroids.f32 -> nearest_centroids.i64/.f32 [nq][nprobe]
//   pf_client_check rank    <dir>   params.txt (K = coarse_probe field), precise_scores.f32 [nq][K], coarse_ids.i64 [nq][K],
//                                   groundtruth.i32 [nq][gt_k] -> ranked.i64 [nq][K], benchmark.txt
//   pf_client_check coarse-rank <dir>  scores.f32, labels.i64, list_sizes.u64 -> coarse_ranked.i64 / .f32 (packed per query)
//   pf_client_check request <dir>   queries_seeded.bin/.off, nearest_idx.i64 [nq][nprobe] -> request.json (the POST body)
//   pf_client_check respond <dir>   response.json (the endpoint's answer) -> scores.f32, list_sizes.u64, labels_out.i64, budget.txt
//   pf_client_check decrypt <dir>   + results.bin/.off, probed_sizes.u64, results_per_query.u64, labels.i64
//                                   -> scores.f32, list_sizes.u64, nearest.i64, budget.txt
// params.txt: dim N t m g nq nprobe coarse_probe k p_0 ... p_{k-1}
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>

#include "pf_client.hpp"

namespace {
template <class T>
std::vector<T> read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<T> out(raw.size() / sizeof(T));
    memcpy(out.data(), raw.data(), out.size() * sizeof(T));
    return out;
}
template <class T>
void write_file(const std::string &path, const std::vector<T> &v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
    if (!f) throw std::runtime_error("cannot write " + path);
}
} // namespace

int main(int argc, char **argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: pf_client_check keygen|nearest|coarse-rank|rank|request|respond|decrypt <dir>\n");
        return 2;
    }
    try {
        const std::string mode = argv[1], dir = std::string(argv[2]) + "/";
        std::ifstream pf(dir + "params.txt");
        uint64_t dim, N, t, m, g, nq, nprobe, coarse_probe, k;
        if (!(pf >> dim >> N >> t >> m >> g >> nq >> nprobe >> coarse_probe >> k)) throw std::runtime_error("params.txt");
        std::vector<uint64_t> primes(k);
        for (auto &p : primes)
            if (!(pf >> p)) throw std::runtime_error("params.txt: primes");
        if (mode == "nearest") { // stage 1 needs no keys
            const std::vector<float> qf = read_file<float>(dir + "queries.f32"), cf = read_file<float>(dir + "centroids.f32");
            if (qf.size() != nq * dim || cf.size() % dim || cf.size() / dim < nprobe) throw std::runtime_error("queries.f32 / centroids.f32 size");
            const auto nearest = prefhetch::Client::sort_nearest_centroids(qf.data(), nq, cf.data(), cf.size() / dim, (uint32_t)dim);
            std::vector<int64_t> idx;
            std::vector<float> dist;
            for (const auto &q : nearest)
                for (uint64_t j = 0; j < nprobe; j++) {
                    idx.push_back(q[j].idx);
                    dist.push_back(q[j].distance);
                }
            write_file(dir + "nearest_centroids.i64", idx);
            write_file(dir + "nearest_centroids.f32", dist);
            printf("ok nearest: %llu queries x %llu of %zu centroids\n", (unsigned long long)nq, (unsigned long long)nprobe, cf.size() / dim);
            return 0;
        }
        if (mode == "coarse-rank") { // compute_nearest_coarse_vectors on a plaintext response
            const std::vector<float> sc = read_file<float>(dir + "scores.f32");
            const std::vector<int64_t> lb = read_file<int64_t>(dir + "labels.i64");
            const std::vector<uint64_t> ls = read_file<uint64_t>(dir + "list_sizes.u64");
            const auto nearest = prefhetch::Client::compute_nearest_coarse_vectors(sc, lb, ls, coarse_probe);
            std::vector<int64_t> oi;
            std::vector<float> od;
            for (const auto &q : nearest)
                for (const auto &e : q) {
                    oi.push_back(e.idx);
                    od.push_back(e.distance);
                }
            write_file(dir + "coarse_ranked.i64", oi);
            write_file(dir + "coarse_ranked.f32", od);
            printf("ok coarse-rank\n");
            return 0;
        }
        if (mode == "rank") { // stage 3 + the reference's recall bookkeeping
            const std::vector<float> ps = read_file<float>(dir + "precise_scores.f32");
            const std::vector<int64_t> cid = read_file<int64_t>(dir + "coarse_ids.i64");
            const bool have_gt = std::ifstream(dir + "groundtruth.i32").good();     // without it: the ranking only
            const std::vector<int32_t> gt = have_gt ? read_file<int32_t>(dir + "groundtruth.i32") : std::vector<int32_t>();
            const uint64_t K = coarse_probe;
            if (ps.size() != nq * K || cid.size() != nq * K || gt.size() % nq) throw std::runtime_error("rank inputs");
            std::vector<std::vector<prefhetch::DistanceIndexData>> coarse(nq);
            for (uint64_t i = 0; i < nq; i++)
                for (uint64_t j = 0; j < K; j++) coarse[i].push_back({0.0f, cid[i * K + j]});
            const auto ranked = prefhetch::Client::compute_nearest_precise_vectors(ps.data(), coarse, K);
            std::vector<int64_t> out;
            for (const auto &q : ranked)
                for (const auto &e : q) out.push_back(e.idx);
            write_file(dir + "ranked.i64", out);
            std::vector<float> outd;
            for (const auto &q : ranked)
                for (const auto &e : q) outd.push_back(e.distance);
            write_file(dir + "ranked.f32", outd);
            if (have_gt) {
                const auto b = prefhetch::Client::benchmark_results(out.data(), nq, K, gt.data(), gt.size() / nq);
                std::ofstream(dir + "benchmark.txt") << b.recall_1 << " " << b.recall_10 << " " << b.recall_100 << " " << b.mrr_1 << " " << b.mrr_10
                                                     << " " << b.mrr_100 << "\n";
            }
            printf("ok rank\n");
            return 0;
        }
        const std::vector<uint8_t> seed_v = read_file<uint8_t>(dir + "seed.bin");
        const std::vector<int64_t> queries = read_file<int64_t>(dir + "queries.i64");
        if (seed_v.size() != 64 || queries.size() != nq * dim) throw std::runtime_error("seed.bin / queries.i64 size");
        std::array<uint8_t, 64> seed;
        memcpy(seed.data(), seed_v.data(), 64);
        prefhetch::Client cl((uint32_t)dim, N, primes, t, (uint32_t)m, (uint32_t)g);
        cl.generateKeys(seed);
        if (mode == "keygen") {
            write_file(dir + "sk.i8", std::vector<int8_t>(cl.secretKeyCoefficients()));
            {   // the same key set twice: seeded (what travels) and full, from two clients with the same seed
                prefhetch::Client twin((uint32_t)dim, N, primes, t, (uint32_t)m, (uint32_t)g);
                twin.generateKeys(seed);
                write_file(dir + "galois_keys_full.bin", twin.galoisKeys(false));
            }
            write_file(dir + "galois_keys.bin", cl.galoisKeys(true));
            for (int seeded = 1; seeded >= 0; seeded--) {
                std::vector<uint8_t> blob;
                std::vector<uint64_t> offs{0};
                for (uint64_t i = 0; i < nq; i++) {
                    std::vector<uint64_t> o;
                    const std::vector<uint8_t> b = cl.compute_encrypted_coarse_query(queries.data() + i * dim, &o, seeded != 0);
                    for (size_t a = 1; a < o.size(); a++) offs.push_back(blob.size() + o[a]);
                    blob.insert(blob.end(), b.begin(), b.end());
                }
                const std::string stem = dir + (seeded ? "queries_seeded" : "queries_full");
                write_file(stem + ".bin", blob);
                write_file(stem + ".off", offs);
            }
            // BatchEncoder probe: encode(0, 1, 2, ...) so the test can compare the plaintext polynomial itself
            std::vector<prefhetch::u64> ramp(N);
            for (uint64_t i = 0; i < N; i++) ramp[i] = (i * 2654435761ULL + 17) % t;
            const std::vector<prefhetch::u64> plain = cl.encode(ramp);
            if (cl.decode(plain) != ramp) throw std::runtime_error("encode / decode do not invert each other");
            write_file(dir + "encode_probe.u64", std::vector<uint64_t>(plain.begin(), plain.end()));
            printf("ok keygen: %llu queries x %u ciphertexts, %u rotation keys, %u candidates per result\n", (unsigned long long)nq,
                   cl.queryCiphertexts(), cl.rotations() - 1, cl.candidatesPerResult());
            return 0;
        }
        if (mode == "request") {
            const std::vector<uint8_t> blob = read_file<uint8_t>(dir + "queries_seeded.bin");
            const std::vector<uint64_t> offs64 = read_file<uint64_t>(dir + "queries_seeded.off");
            const std::vector<int64_t> nearest = read_file<int64_t>(dir + "nearest_idx.i64");
            if (nearest.size() != nq * nprobe) throw std::runtime_error("nearest_idx.i64 size");
            const std::string body = prefhetch::Client::coarse_search_encrypted_request(blob, offs64, nearest.data(), nq, nprobe);
            std::ofstream(dir + "request.json", std::ios::binary) << body;
            printf("ok request: %zu bytes\n", body.size());
            return 0;
        }
        if (mode == "respond") {
            std::ifstream f(dir + "response.json", std::ios::binary);
            const std::string body((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            const prefhetch::EncryptedCoarseResponse resp = prefhetch::Client::parse_coarse_search_encrypted_response(body);
            std::vector<float> scores;
            std::vector<int64_t> labels;
            std::vector<uint64_t> list_sizes;
            int budget = 0;
            cl.get_coarse_scores(resp, queries.data(), scores, labels, list_sizes, &budget);
            write_file(dir + "scores.f32", scores);
            write_file(dir + "list_sizes.u64", list_sizes);
            write_file(dir + "labels_out.i64", labels);
            std::ofstream(dir + "budget.txt") << budget << "\n";
            printf("ok respond: %zu scores, min noise budget %d bits\n", scores.size(), budget);
            return 0;
        }
        if (mode == "decrypt") {
            const std::vector<uint8_t> results = read_file<uint8_t>(dir + "results.bin");
            const std::vector<uint64_t> roff = read_file<uint64_t>(dir + "results.off");
            const std::vector<uint64_t> probed = read_file<uint64_t>(dir + "probed_sizes.u64");
            const std::vector<uint64_t> rpq = read_file<uint64_t>(dir + "results_per_query.u64");
            const std::vector<int64_t> labels = read_file<int64_t>(dir + "labels.i64");
            if (probed.size() != nq * nprobe || rpq.size() != nq || roff.empty()) throw std::runtime_error("response envelope sizes");
            std::vector<uint64_t> list_sizes;
            int budget = 0;
            const std::vector<float> scores = cl.decrypt_coarse_scores(
                nq, (uint32_t)nprobe, queries.data(), probed.data(), rpq.data(),
                [&](uint64_t r) {
                    if (r + 1 >= roff.size() || roff[r + 1] > results.size() || roff[r] > roff[r + 1]) throw std::runtime_error("result index beyond the response");
                    return std::pair<const uint8_t *, size_t>(results.data() + roff[r], (size_t)(roff[r + 1] - roff[r]));
                },
                &list_sizes, &budget);
            if (labels.size() != scores.size()) throw std::runtime_error("labels and scores differ in length");
            const auto nearest = prefhetch::Client::compute_nearest_coarse_vectors(scores, labels, list_sizes, coarse_probe);
            std::vector<int64_t> top;
            for (const auto &q : nearest)
                for (uint64_t j = 0; j < coarse_probe; j++) top.push_back(q[j].idx);
            write_file(dir + "scores.f32", scores);
            write_file(dir + "list_sizes.u64", list_sizes);
            write_file(dir + "nearest.i64", top);
            std::ofstream(dir + "budget.txt") << budget << "\n";
            printf("ok decrypt: %zu scores, min noise budget %d bits\n", scores.size(), budget);
            return 0;
        }
        fprintf(stderr, "unknown mode %s\n", mode.c_str());
        return 2;
    } catch (const std::exception &ex) {
        fprintf(stderr, "pf_client_check: %s\n", ex.what());
        return 1;
    }
}
