/*
 * prefhetch_b200.h — C ABI of the B200-native engine for PreFHEtch's server-side search hot path.
 *
 * The reference has no plugin/FFI layer: its seam is the `Server` member-function surface
 * (ref: include/server/server_lib.h:25-49) called from the Drogon handlers
 * (ref: src/server/controllers/Query.cc:18,49,87).  Each entry point below names the reference
 * interface it replaces.  All pointers are HOST pointers unless the name says `_device`; shapes
 * are run-time values (the reference's are compile-time constants,
 * ref: include/common/client_server_utils.h:10-20).  No exceptions cross this boundary: every
 * call returns a status code and pf_last_error() gives the message
 * (the reference throws std::runtime_error, ref: src/server/server_lib.cpp:66,94).
 * Calls on one engine are serialised internally (the reference `Server` is a non-re-entrant
 * singleton, ref: include/server/server_lib.h:20-23, src/server/server_lib.cpp:121).
 *
 * Ciphertext word layout everywhere: uint64 [poly 2][limb L][coeff N] (SEAL Ciphertext::data()).
 * "coefficient form" / "NTT form" are SEAL's (bit-reversed NTT output, smallest primitive root).
 */
#ifndef PREFHETCH_B200_H
#define PREFHETCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 2
#define PF_MAX_PRIMES 16

enum {
    PF_OK = 0,
    PF_ERR_INVALID = 1,  /* bad argument / unsupported parameter set */
    PF_ERR_CUDA = 2,     /* CUDA runtime failure (no CPU fallback exists) */
    PF_ERR_CAPACITY = 3, /* caller buffer too small; required size reported where documented */
    PF_ERR_STATE = 4,    /* call out of order (no index / no Galois key loaded) */
    PF_ERR_FORMAT = 5    /* malformed serialized ciphertext / key */
};

typedef struct pf_engine pf_engine;

typedef struct {
    uint32_t struct_size;   /* sizeof(pf_params) */
    int32_t device;         /* CUDA device ordinal */
    uint64_t poly_degree;   /* N: power of two in [1024, 16384] */
    uint32_t num_primes;    /* k = data primes + 1 special prime (SEAL key-level chain) */
    uint32_t dim;           /* vector dimension d (ref: PRECISE_VECTOR_DIMENSIONS) */
    uint64_t primes[PF_MAX_PRIMES];
    uint64_t plain_modulus; /* t, batching prime */
    uint32_t query_cts;     /* m: ciphertexts per query (dimension chunks), power of two */
    uint32_t partial_g;     /* g: partial sums per candidate, power of two dividing d_pad/m */
    uint32_t rank;          /* this engine keeps the IVF lists l with l % world == rank */
    uint32_t world;
    uint32_t result_limbs;  /* 0 = L.  < L: results are mod-switched down to this many limbs before they
                             * leave the GPU (SEAL Evaluator::mod_switch_to_inplace), shrinking the response */
    uint32_t reserved;
} pf_params;

typedef struct {
    uint64_t nlist, ntotal, nblocks, nblocks_local;
    uint64_t db_bytes;      /* NTT-domain plaintext bytes resident in HBM (this shard) */
    uint32_t K, C, R, d_pad; /* diagonals per block, candidates per block, rotations per chunk */
    uint32_t L, k;
} pf_index_info;

/* ---- life cycle ------------------------------------------------------------------------- */
int pf_abi_version(void);
/* replaces: Server::Server (ref: src/server/server_lib.cpp:32-46).  Fails with PF_ERR_CUDA when no
 * CUDA device is usable: there is deliberately no CPU path. */
int pf_engine_create(const pf_params *params, pf_engine **out);
void pf_engine_destroy(pf_engine *e);
/* last error message of this engine (or of the calling thread when e == NULL) */
const char *pf_last_error(const pf_engine *e);
/* stream all engine work is launched on (a cudaStream_t); set to run on the caller's stream */
void *pf_engine_stream(pf_engine *e);
int pf_engine_set_stream(pf_engine *e, void *cuda_stream);
int pf_engine_synchronize(pf_engine *e);

/* ---- index ------------------------------------------------------------------------------ */
/* replaces: Server::init_index's hand-over of the trained IVF index to the searcher
 * (ref: src/server/server_lib.cpp:55-99): centroids [nlist][d] as `quantizer->reconstruct` returns
 * them (:107), inverted lists CSR style (list_offsets[nlist+1]; ids / vectors in list order, i.e.
 * FAISS invlists order).  Vectors must be integer valued in [0,255] for the encrypted path
 * (SIFT is; ref data set dataset.sh).  Encodes every block of this rank's lists into NTT-domain
 * plaintext diagonals in HBM (kernels, not host code). */
int pf_load_index(pf_engine *e, uint64_t nlist, const float *centroids, const int64_t *list_offsets,
                  const int64_t *ids, const float *vectors);
int pf_get_index_info(pf_engine *e, pf_index_info *out);
/* replaces: Server::retrieve_centroids (ref: src/server/server_lib.cpp:101-109) */
int pf_retrieve_centroids(pf_engine *e, float *out, uint64_t cap_floats);

/* ---- stage 1: coarse quantization ------------------------------------------------------- */
/* replaces: sort_nearest_centroids + the first-NPROBE slice (ref: src/client/client_lib.cpp:50-81,
 * :93-103).  Same arithmetic (float difference, double square, float running sum), ascending,
 * ties by lower index.  out_idx [nq][nprobe]; out_dist optional. */
int pf_coarse_quantize(pf_engine *e, uint64_t nq, const float *x, uint32_t nprobe, int64_t *out_idx,
                       float *out_dist);

/* ---- stage 2, plaintext ------------------------------------------------------------------ */
/* replaces: the m_Index->search_encrypted call in Server::coarseSearch
 * (ref: src/server/server_lib.cpp:111-138): for every query, for each given list in order, one
 * (distance, id) per stored vector, packed back to back; list_sizes[i] = candidates of query i.
 * Distances are the exact squared L2 of Server::preciseSearch (ref: :140-167).  If *total > cap
 * nothing past cap is written and PF_ERR_CAPACITY is returned with *total = required entries. */
int pf_search_lists_plain(pf_engine *e, uint64_t nq, const float *x, const int64_t *idx, uint32_t nprobe,
                          float *dist, int64_t *labels, uint64_t cap, uint64_t *list_sizes, uint64_t *total);
/* The same call with the distance the reference's FAISS fork computes TODAY: product-quantizer ADC
 * (replaces: faiss::IndexIVFPQ::search_encrypted as called at ref: src/server/server_lib.cpp:126-130 on the index
 * built at :34-36 with SUB_QUANTIZERS x SUB_QUANTIZER_SIZE bits, include/common/client_server_utils.h:19-20).
 * pf_load_pq gives the engine the product quantizer of the loaded index: M sub-quantizers of nbits = 8 bits,
 * pq_centroids [M][256][d/M] (FAISS ProductQuantizer::centroids as stored in the .faiss file) and codes
 * [ntotal][M], one row per vector in the order of pf_load_index's `ids` (the invlists' code arrays back to back).
 * pf_search_lists_pq packs (distance, id) exactly like pf_search_lists_plain; distance = sum over sub-quantizers of
 * ||(x - centroid[l])_m - pq[m][code_m]||^2 in float (by_residual, METRIC_L2).  [EXT]: restated from the published
 * FAISS algorithm (the fork's source is absent); equal to a FAISS build up to float rounding, not bit for bit. */
int pf_load_pq(pf_engine *e, uint32_t M, uint32_t nbits, const float *pq_centroids, const uint8_t *codes);
int pf_search_lists_pq(pf_engine *e, uint64_t nq, const float *x, const int64_t *idx, uint32_t nprobe,
                       float *dist, int64_t *labels, uint64_t cap, uint64_t *list_sizes, uint64_t *total);
/* replaces: Server::preciseSearch (ref: src/server/server_lib.cpp:140-167): ids [nq][nids] are
 * base-file row numbers, out [nq][nids]. */
int pf_precise_search(pf_engine *e, uint64_t nq, const float *x, const int64_t *ids, uint32_t nids, float *out);

/* ---- stage 2, encrypted ------------------------------------------------------------------ */
/* Galois key for element `galois_elt`, words [L][2][k][N] in NTT form (SEAL KSwitchKeys data of
 * that element, GaloisKeys::key(galois_elt)). */
int pf_set_galois_key(pf_engine *e, uint32_t galois_elt, const uint64_t *key_words);
/* SEAL-serialized GaloisKeys (compr_mode none), all elements it holds */
int pf_load_galois_keys(pf_engine *e, const uint8_t *bytes, size_t len);
uint32_t pf_galois_elt_from_step(pf_engine *e, int step);

typedef struct {
    uint64_t nresults;        /* result ciphertexts written (sum over queries) */
    uint64_t out_bytes;       /* bytes written to out_cts */
    uint64_t useful_distances; /* real candidates covered */
    uint64_t slot_distances;   /* candidates incl. padding = nresults * C */
} pf_search_stats;

/* The encrypted variant of Server::coarseSearch (additive to ref: src/server/controllers/Query.cc:29-63):
 * query_cts[0, query_bytes) holds nq*m SEAL-serialized BFV ciphertexts (coefficient form, top level),
 * ct_offsets[nq*m+1] their byte offsets (ascending, all <= query_bytes, else PF_ERR_INVALID).  A ciphertext whose
 * parms_id is neither this engine's top data level (pf_parms_id over the L data primes) nor all zero ("not
 * stamped") is refused with PF_ERR_FORMAT, as seal::Ciphertext::load(context, ...) refuses it.  For query i
 * and each of its lists idx[i][p] (in order, lists not owned by this rank are skipped) one result
 * ciphertext per block of the list is written to out_cts in SEAL format (coefficient form).  Result r is
 * the pf_result_serialized_size() bytes starting at result_offsets[r]; results sit in slots of
 * pf_result_slot_size() bytes so that their words are 128-byte aligned for the device-to-host copy
 * (out_cap >= nresults * slot; result_offsets[nresults] = bytes used).  Pinned, 128-byte aligned out_cts
 * gives the fastest copies.  results_per_query[nq].  labels / list_sizes as in pf_search_lists_plain (ids
 * of the owned probed lists, packed); probed_sizes[nq][nprobe] = length of each probed list (0 when not
 * owned) so the client can map candidate j of a list to (result j / C, candidate j % C). */
int pf_search_lists_encrypted(pf_engine *e, uint64_t nq, const uint8_t *query_cts, uint64_t query_bytes,
                              const uint64_t *ct_offsets, const int64_t *idx, uint32_t nprobe, uint8_t *out_cts,
                              uint64_t out_cap, uint64_t *result_offsets, uint64_t max_results,
                              uint64_t *results_per_query, int64_t *labels, uint64_t label_cap, uint64_t *list_sizes,
                              uint64_t *probed_sizes, pf_search_stats *stats);

/* The same call split in two, for a handler thread that keeps the GPU busy across requests
 * (ref: Query::coarse_search is invoked per request, src/server/controllers/Query.cc:29-63):
 * pf_search_submit validates, plans, enqueues upload + compute + download and returns; every host-side
 * output except out_cts (offsets, sizes, labels, stats) is final on return.  pf_search_collect waits until
 * the result ciphertexts of that ticket are in out_cts.  Up to 4 searches may be in flight per engine
 * (a fifth submit returns PF_ERR_STATE); with three, request i+2 is queued behind the compute of request
 * i+1 while request i is downloaded, and neither the GPU nor the PCIe link waits for the host.  query_cts and out_cts must stay valid and untouched until collect returns. */
int pf_search_submit(pf_engine *e, uint64_t nq, const uint8_t *query_cts, uint64_t query_bytes,
                     const uint64_t *ct_offsets, const int64_t *idx, uint32_t nprobe, uint8_t *out_cts,
                     uint64_t out_cap, uint64_t *result_offsets, uint64_t max_results, uint64_t *results_per_query,
                     int64_t *labels, uint64_t label_cap, uint64_t *list_sizes, uint64_t *probed_sizes,
                     pf_search_stats *stats, uint64_t *ticket);
int pf_search_collect(pf_engine *e, uint64_t ticket);
/* query groups a search is cut into (copy/compute overlap INSIDE one call): 0 = default (4, best latency
 * for a lone call); 1 = whole-batch kernels, best throughput when calls are pipelined with submit/collect */
int pf_search_set_groups(pf_engine *e, uint32_t groups);
/* page-lock caller memory (e.g. a POSIX shared-memory response buffer that every rank of a node writes its
 * share of the response into) so that copies to / from it are asynchronous DMA */
int pf_host_register(pf_engine *e, void *ptr, size_t bytes);
int pf_host_unregister(pf_engine *e, void *ptr);

/* Device-resident form of the same step (what `value` in bench.py times): d_query_cts is a DEVICE
 * pointer to raw words [nq][m][2][L][N] (coefficient form); d_out a DEVICE buffer of
 * cap_results*2*result_limbs*N words receiving coefficient-form results in the order above.  idx is a host
 * array.  Asynchronous on the engine stream. */
int pf_search_device(pf_engine *e, uint64_t nq, const uint64_t *d_query_cts, const int64_t *idx, uint32_t nprobe,
                     uint64_t *d_out, uint64_t cap_results, uint64_t *results_per_query, pf_search_stats *stats);

/* ---- multi-GPU: peer-memory gather of result ciphertexts ---------------------------------------
 * One process per GPU.  Rank 0 allocates one gather buffer per peer with pf_ipc_alloc and ships the 64-byte
 * handles to the other processes (any host channel); they map theirs with pf_ipc_open and, after every step,
 * copy their result ciphertexts into it with pf_copy_async on a side stream (copy-engine DMA over NVLink,
 * overlapped with the next step: no SM is used and no collective queues behind the compute CTAs), then raise
 * their arrival flag in rank 0's memory with pf_flag_write; rank 0 waits for the flags with pf_flag_wait and
 * acknowledges.  (The mapped pointer may also be passed as d_out of pf_search_device, so that the last kernel
 * of the step stores straight into rank 0's HBM; measured slower at 8 GPUs — seven writers burst at once.) */
#define PF_IPC_HANDLE_BYTES 64
int pf_ipc_alloc(pf_engine *e, size_t bytes, void **dptr, uint8_t handle[PF_IPC_HANDLE_BYTES]);
int pf_ipc_open(pf_engine *e, const uint8_t handle[PF_IPC_HANDLE_BYTES], void **dptr);
int pf_ipc_close(pf_engine *e, void *dptr);
int pf_ipc_free(pf_engine *e, void *dptr);
/* Stream-ordered flags in (peer-mapped) device memory: one-thread kernels, so they slip in beside the
 * compute kernels where an NCCL collective would queue behind them.  write: *flag = value after all
 * prior work of the stream (system-scope release).  wait: the stream blocks until *flag >= value.
 * cuda_stream NULL = the engine stream.  Waits must only target flags whose writer does not itself
 * wait on this stream's later work (the bench protocol: arrival flags -> rank 0 -> ack flags). */
/* copy-engine (DMA) device-to-device copy, e.g. local results -> peer-mapped gather buffer, on a
 * caller stream so that it overlaps the next step's kernels without using SMs */
int pf_copy_async(pf_engine *e, void *dst, const void *src, size_t bytes, void *cuda_stream);
int pf_flag_write(pf_engine *e, void *flag, uint32_t value, void *cuda_stream);
/* The wait is bounded (20 s, env PF_FLAG_TIMEOUT_MS): if the flag never arrives the kernel gives up, the
 * stream continues, and this and every later call on the engine returns PF_ERR_CUDA — a dead peer does
 * not hang the process. */
int pf_flag_wait(pf_engine *e, const void *flag, uint32_t value, void *cuda_stream);
/* position-weighted 64-bit checksum of nwords device words (verifies that what landed in a gather buffer is
 * what the producing rank computed); synchronous on cuda_stream (NULL = engine stream) */
int pf_device_checksum(pf_engine *e, const void *dptr, uint64_t nwords, uint64_t *out, void *cuda_stream);

/* per-phase device timers (CUDA events on the engine stream), accumulated since the last reset */
enum { PF_T_COARSE = 0, PF_T_TONTT = 1, PF_T_ROTATE = 2, PF_T_MAC = 3, PF_T_INTT = 4, PF_T_COUNT = 8 };
int pf_timing_enable(pf_engine *e, int on);
int pf_timing_read(pf_engine *e, float *ms /*[PF_T_COUNT]*/, uint64_t *launches /*[PF_T_COUNT]*/, int reset);
/* kernels launched by this engine since creation (all phases) */
uint64_t pf_launch_count(pf_engine *e);

/* ---- primitives (parity-test entry points; host buffers, results copied back) ------------ */
/* in-place forward / inverse negacyclic NTT of npoly polynomials; limb[i] selects the modulus of
 * polynomial i (0..k-1 coefficient primes, -1 = plain modulus t).  SEAL: util/ntt.cpp. */
int pf_ntt_forward(pf_engine *e, uint64_t *polys, uint64_t npoly, const int32_t *limb);
int pf_ntt_inverse(pf_engine *e, uint64_t *polys, uint64_t npoly, const int32_t *limb);
/* out[2][L][N] = sum_{k<K} cts[k] (.) pts[k]  (+ addend[L][N] on polynomial 0 when non-NULL);
 * SEAL: Evaluator::multiply_plain (NTT) + add_inplace chain. */
int pf_ct_pt_mac(pf_engine *e, const uint64_t *cts, const uint64_t *pts, uint32_t K, const uint64_t *addend,
                 uint64_t *out);
/* SEAL: Evaluator::add */
int pf_ct_add(pf_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* SEAL: Evaluator::transform_to_ntt_inplace / transform_from_ntt_inplace on ncts ciphertexts */
int pf_ct_to_ntt(pf_engine *e, uint64_t *cts, uint64_t ncts);
int pf_ct_from_ntt(pf_engine *e, uint64_t *cts, uint64_t ncts);
/* SEAL: Evaluator::rotate_rows on a coefficient-form ciphertext (key for 3^step must be loaded) */
int pf_rotate_rows(pf_engine *e, const uint64_t *ct, int step, uint64_t *out);
/* rotated query set: cts[m][2][L][N] coefficient form -> rot[K][2][L][N] NTT form.  chain != 0 applies
 * the step-1 key repeatedly, otherwise the key of every step 1..R-1 is used on the input. */
int pf_rotate_query_set(pf_engine *e, const uint64_t *cts, int chain, uint64_t *rot);
/* SEAL: BatchEncoder::encode of N slot values (mod t) */
int pf_batch_encode(pf_engine *e, const uint64_t *values, uint64_t *plain);
/* encode one block of nvec integer vectors xs[nvec][d]: diag[K][L][N], norm[L][N] (NTT form) */
int pf_encode_block(pf_engine *e, const int32_t *xs, uint32_t nvec, uint64_t *diag, uint64_t *norm);
/* SEAL wire format (compr_mode none) of a coefficient-form size-2 ciphertext */
size_t pf_ct_serialized_size(pf_engine *e);
/* bytes of out_cts reserved per result ciphertext by pf_search_lists_encrypted, and the length of a
 * result's SEAL stream (results have result_limbs limbs) */
size_t pf_result_slot_size(pf_engine *e);
size_t pf_result_serialized_size(pf_engine *e);
/* Normalises one SEAL stream (16-byte SEALHeader + body) to compr_mode none: zlib- and zstd-compressed streams
 * are inflated (zstd through libzstd.so.1, bound at run time: without it such a stream is a format error),
 * uncompressed ones copied; corrupt input or an unknown compr_mode -> PF_ERR_FORMAT; too small a buffer ->
 * PF_ERR_CAPACITY with *written = bytes needed.  (What seal::Serialization::Load does before the members
 * are read; [EXT] SEAL 4.1 serialization.cpp.)  Needs no engine and no GPU. */
int pf_seal_stream_inflate(const uint8_t *in, size_t len, uint8_t *out, size_t cap, size_t *written, size_t *consumed);
/* Seeded ciphertexts (seal::Serializable<Ciphertext> of a symmetric-key Encryptor: c1 is replaced by the seed of
 * the PRNG that drew it; [EXT] SEAL 4.1 Ciphertext::save_members / expand_seed, Blake2xbPRNG,
 * sample_poly_uniform): writes the equivalent full stream (compr_mode none) of `in` — c1 re-created over the
 * given data primes — or a copy when `in` is not seeded; zlib / zstd input is inflated first.  pf_search_submit /
 * pf_search_lists_encrypted / pf_ct_deserialize accept seeded streams directly (expanded on the host before
 * the upload: a slow path like zlib).  PF_ERR_FORMAT for malformed input or a PRNG other than blake2xb;
 * PF_ERR_CAPACITY with *written = bytes needed.  Needs no engine and no GPU. */
int pf_seal_ct_expand(const uint8_t *in, size_t len, uint64_t poly_degree, const uint64_t *data_primes, uint32_t nprimes,
                      uint8_t *out, size_t cap, size_t *written, size_t *consumed);
/* The same expansion ON THE DEVICE, as the search calls do it when every query stream of a request is an
 * uncompressed blake2xb-seeded one (half the upload of a symmetric-key SEAL client): `in` is one such stream of
 * this engine's top level, ct_words receives [2][L][N] (c0 as sent, c1 = sample_poly_uniform of the seeded
 * Blake2xb PRNG: one CTA per 4096-byte PRNG refill, rejected words re-drawn in stream order).  Bit-identical to
 * pf_seal_ct_expand.  PF_SEEDED_HOST=1 makes the search calls expand on the host instead. */
int pf_seal_ct_expand_device(pf_engine *e, const uint8_t *in, size_t len, uint64_t *ct_words, size_t cap_words);
/* pf_seal_ct_expand for a whole request: the ncts streams in[offsets[c], offsets[c+1]) -> their full compr_mode none
 * forms back to back in out (out_offsets[ncts + 1]; PF_ERR_CAPACITY with out_offsets filled when out is too small).
 * This is the slow path pf_search_submit runs on compressed / seeded queries: the streams are independent and are
 * handled by up to `threads` host threads (0 = the engine's default: hardware threads, at most 16, env
 * PF_HOST_THREADS).  Needs no engine and no GPU. */
int pf_seal_ct_expand_batch(const uint8_t *in, size_t in_bytes, const uint64_t *offsets, uint64_t ncts, uint64_t poly_degree,
                            const uint64_t *data_primes, uint32_t nprimes, uint8_t *out, size_t cap, uint64_t *out_offsets,
                            uint32_t threads);
/* The same for keys: a SEAL GaloisKeys stream as seal::Serializable<GaloisKeys> saves it (every key ciphertext seeded:
 * c1 replaced by its PRNG seed — what KeyGenerator::create_galois_keys returns without a destination; [EXT] SEAL 4.1
 * keygenerator.cpp, kswitchkeys.h) -> the equivalent full compr_mode none stream over the k = L + 1 key primes;
 * full streams are copied, zlib / zstd inflated.  pf_load_galois_keys accepts seeded streams directly; this entry
 * point needs no engine and no GPU.  PF_ERR_CAPACITY with *written = bytes needed when out is too small. */
int pf_seal_galois_keys_expand(const uint8_t *in, size_t len, uint64_t poly_degree, const uint64_t *key_primes, uint32_t nprimes,
                               uint8_t *out, size_t cap, size_t *written);
/* SEAL parms_id of the BFV parameter set {poly_degree, coeff_primes[0..nprimes), plain_modulus}: BLAKE2b-256
 * of {scheme = 1, N, primes..., t} as 4 little-endian words (replaces EncryptionParameters::parms_id();
 * [EXT] SEAL 4.1 encryptionparams.cpp compute_parms_id).  Needs no engine and no GPU. */
int pf_parms_id(uint64_t poly_degree, const uint64_t *coeff_primes, uint32_t nprimes, uint64_t plain_modulus,
                uint64_t out[4]);
/* parms_id written into result ciphertexts (SEAL: context_data(level)->parms_id()); defaults to the
 * query's parms_id when result_limbs == L, else pf_parms_id of the first result_limbs data primes */
int pf_set_result_parms_id(pf_engine *e, const uint64_t parms_id[4]);
int pf_ct_serialize(pf_engine *e, const uint64_t *ct, int is_ntt, uint8_t *out, size_t cap, size_t *written);
/* accepts ciphertexts with L (query level) or result_limbs limbs; *limbs receives the count */
int pf_ct_deserialize(pf_engine *e, const uint8_t *in, size_t len, uint64_t *ct, size_t cap_words, int *limbs,
                      int *is_ntt, size_t *consumed);

#ifdef __cplusplus
}
#endif
#endif
