// tests/host_standin/standin.cpp — TEST INFRASTRUCTURE, never part of the product.
//
// An oracle-backed stand-in for the subset of the C ABI that the host C++ layer calls (prefhetch::Server in
// host/pf_server.hpp, the handler bodies, pf_server_check, pf_roundtrip_example), so that THAT layer — argument
// marshalling, buffer sizing, the JSON envelope, offsets, the client — can be exercised end to end on a machine
// without a GPU (tests/test_host_cpp_standin.py builds it into the test's tmp directory and links the check
// programs against it instead of libprefhetch_b200.so).  The product library has no CPU path and nothing in
// prefhetch_b200/ knows this file exists; the arithmetic here is the oracle's (tests/ may link oracle/), the SEAL
// stream normalisation is the product's own engine-less host code (pf_seal_ct_expand_batch,
// pf_seal_galois_keys_expand, pf_parms_id), loaded from the real library with dlopen.
// It found the dangling string_views fixed in commit "Fix dangling string_views in the C++ checks".
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "prefhetch_b200.h"
extern "C" {
#include "pf_oracle.h"
}
struct pf_engine {
    pf_params p;
    pfo_context *ctx = nullptr;
    pfo_layout lay;
    int L, k, Lr;
    uint64_t N;
    std::vector<float> cent, vecs;
    std::vector<int64_t> offs, ids;
    std::vector<long long> pos_of_id;
    uint64_t nlist = 0, ntotal = 0;
    std::vector<int64_t> block_off;
    std::vector<uint32_t> block_n;
    std::vector<size_t> list_first_block;
    std::vector<uint64_t> diag, norm;
    std::map<uint32_t, std::vector<uint64_t>> keys;
    std::string err;
};
static std::string g_err;
static void *real() {
    static void *h = dlopen(getenv("PF_STANDIN_REAL_LIB"), RTLD_NOW | RTLD_LOCAL); // the product library, for its host-only parsers
    if (!h) throw std::runtime_error(dlerror());
    return h;
}
template <class F> F sym(const char *n) { return reinterpret_cast<F>(dlsym(real(), n)); }
static int fail(pf_engine *e, int code, const std::string &m) { (e ? e->err : g_err) = m; return code; }
extern "C" {
int pf_engine_create(const pf_params *p, pf_engine **out) {
    auto *e = new pf_engine;
    e->p = *p;
    e->N = p->poly_degree; e->k = (int)p->num_primes; e->L = e->k - 1; e->Lr = p->result_limbs ? (int)p->result_limbs : e->L;
    e->ctx = pfo_context_create(p->poly_degree, p->primes, e->k, p->plain_modulus);
    if (!e->ctx || pfo_layout_init(&e->lay, p->poly_degree, p->dim, p->query_cts, p->partial_g)) { g_err = "bad params"; delete e; return PF_ERR_INVALID; }
    *out = e; return PF_OK;
}
void pf_engine_destroy(pf_engine *e) { if (e) { pfo_context_destroy(e->ctx); delete e; } }
const char *pf_last_error(const pf_engine *e) { return e ? e->err.c_str() : g_err.c_str(); }
int pf_load_index(pf_engine *e, uint64_t nlist, const float *c, const int64_t *lo, const int64_t *ids, const float *v) {
    const uint32_t d = e->p.dim;
    e->nlist = nlist; e->ntotal = (uint64_t)lo[nlist];
    e->cent.assign(c, c + nlist * d); e->offs.assign(lo, lo + nlist + 1); e->ids.assign(ids, ids + e->ntotal); e->vecs.assign(v, v + e->ntotal * d);
    e->pos_of_id.assign(e->ntotal, -1);
    for (uint64_t i = 0; i < e->ntotal; i++) if (ids[i] >= 0 && (uint64_t)ids[i] < e->ntotal) e->pos_of_id[ids[i]] = (long long)i;
    e->block_off.clear(); e->block_n.clear(); e->list_first_block.assign(nlist + 1, 0);
    for (uint64_t l = 0; l < nlist; l++) {
        e->list_first_block[l] = e->block_off.size();
        for (int64_t o = lo[l]; o < lo[l + 1]; o += e->lay.C) { e->block_off.push_back(o); e->block_n.push_back((uint32_t)std::min<int64_t>(e->lay.C, lo[l + 1] - o)); }
    }
    e->list_first_block[nlist] = e->block_off.size();
    std::vector<int32_t> xs(e->ntotal * d);
    for (size_t i = 0; i < xs.size(); i++) xs[i] = (int32_t)v[i];
    const size_t nb = e->block_off.size();
    e->diag.assign(nb * e->lay.K * e->L * e->N, 0); e->norm.assign(nb * e->L * e->N, 0);
    pfo_encode_blocks(e->ctx, &e->lay, nb, xs.data(), e->block_off.data(), e->block_n.data(), e->diag.data(), e->norm.data(), 8);
    return PF_OK;
}
int pf_get_index_info(pf_engine *e, pf_index_info *o) {
    memset(o, 0, sizeof(*o)); o->nlist = e->nlist; o->ntotal = e->ntotal; o->nblocks = o->nblocks_local = e->block_off.size();
    o->K = e->lay.K; o->C = e->lay.C; o->R = e->lay.R; o->d_pad = e->lay.d_pad; o->L = e->L; o->k = e->k; return PF_OK;
}
int pf_retrieve_centroids(pf_engine *e, float *out, uint64_t cap) { if (cap < e->cent.size()) return fail(e, PF_ERR_CAPACITY, "cap"); memcpy(out, e->cent.data(), e->cent.size() * 4); return PF_OK; }
int pf_coarse_quantize(pf_engine *e, uint64_t nq, const float *x, uint32_t nprobe, int64_t *idx, float *dist) {
    std::vector<float> d(nq * nprobe);
    pfo_coarse_quantize(nq, e->p.dim, e->nlist, x, e->cent.data(), nprobe, idx, dist ? dist : d.data()); return PF_OK;
}
int pf_search_lists_plain(pf_engine *e, uint64_t nq, const float *x, const int64_t *idx, uint32_t nprobe, float *dist, int64_t *labels, uint64_t cap,
                          uint64_t *list_sizes, uint64_t *total) {
    std::vector<size_t> ls(nq);
    size_t need = 0;
    for (uint64_t i = 0; i < nq * nprobe; i++)
        if (idx[i] < 0 || (uint64_t)idx[i] >= e->nlist) return fail(e, PF_ERR_INVALID, "list id out of range");
    for (uint64_t i = 0; i < nq * nprobe; i++) need += (size_t)(e->offs[idx[i] + 1] - e->offs[idx[i]]);
    std::vector<float> dt(need + 1); std::vector<int64_t> lb(need + 1);
    size_t w = pfo_search_lists_plain(nq, e->p.dim, x, idx, nprobe, e->offs.data(), e->ids.data(), e->vecs.data(), dt.data(), lb.data(), need, ls.data());
    for (uint64_t i = 0; i < nq; i++) list_sizes[i] = ls[i];
    *total = w;
    if (w > cap) return PF_ERR_CAPACITY;
    memcpy(dist, dt.data(), w * 4); memcpy(labels, lb.data(), w * 8); return PF_OK;
}
int pf_precise_search(pf_engine *e, uint64_t nq, const float *x, const int64_t *ids, uint32_t nids, float *out) {
    const uint32_t d = e->p.dim;
    for (uint64_t i = 0; i < nq; i++) for (uint32_t j = 0; j < nids; j++) {
        const long long pos = e->pos_of_id[ids[i * nids + j]];
        out[i * nids + j] = pfo_l2sqr_ref(e->vecs.data() + (size_t)pos * d, x + i * d, d);
    }
    return PF_OK;
}
uint32_t pf_galois_elt_from_step(pf_engine *e, int step) { return pfo_galois_elt_from_step(e->ctx, step); }
int pf_set_galois_key(pf_engine *e, uint32_t elt, const uint64_t *w) { e->keys[elt].assign(w, w + (size_t)e->L * 2 * e->k * e->N); return PF_OK; }
int pf_load_galois_keys(pf_engine *e, const uint8_t *bytes, size_t len) {
    auto expand = sym<int (*)(const uint8_t *, size_t, uint64_t, const uint64_t *, uint32_t, uint8_t *, size_t, size_t *)>("pf_seal_galois_keys_expand");
    size_t need = 0;
    int rc = expand(bytes, len, e->N, e->p.primes, (uint32_t)e->k, nullptr, 0, &need);
    if (rc != PF_ERR_CAPACITY) return fail(e, PF_ERR_FORMAT, "keys");
    std::vector<uint8_t> full(need);
    if (expand(bytes, len, e->N, e->p.primes, (uint32_t)e->k, full.data(), need, &need)) return fail(e, PF_ERR_FORMAT, "keys2");
    size_t off = 56; const size_t per = 113 + (size_t)2 * e->k * e->N * 8;
    for (uint64_t slot = 0; slot < e->N; slot++) {
        uint64_t dim2; memcpy(&dim2, full.data() + off, 8); off += 8;
        if (!dim2) continue;
        std::vector<uint64_t> w((size_t)e->L * 2 * e->k * e->N);
        for (uint64_t j = 0; j < dim2; j++) { memcpy(w.data() + j * 2 * e->k * e->N, full.data() + off + 113, per - 113); off += per; }
        e->keys[(uint32_t)(2 * slot + 1)] = w;
    }
    return PF_OK;
}
size_t pf_result_slot_size(pf_engine *e) { return 128 + (size_t)2 * e->Lr * e->N * 8; }
size_t pf_result_serialized_size(pf_engine *e) { return 113 + (size_t)2 * e->Lr * e->N * 8; }
int pf_search_submit(pf_engine *e, uint64_t nq, const uint8_t *q, uint64_t qbytes, const uint64_t *coffs, const int64_t *idx, uint32_t nprobe, uint8_t *out,
                     uint64_t out_cap, uint64_t *roffs, uint64_t max_results, uint64_t *rpq, int64_t *labels, uint64_t label_cap, uint64_t *list_sizes,
                     uint64_t *probed, pf_search_stats *st, uint64_t *ticket) {
    const uint32_t m = e->p.query_cts; const size_t ncts = nq * m, ctw = (size_t)2 * e->L * e->N;
    for (size_t c = 0; c < ncts; c++) if (coffs[c + 1] < coffs[c] || coffs[c + 1] > qbytes) return fail(e, PF_ERR_INVALID, "offsets");
    auto batch = sym<int (*)(const uint8_t *, size_t, const uint64_t *, uint64_t, uint64_t, const uint64_t *, uint32_t, uint8_t *, size_t, uint64_t *, uint32_t)>("pf_seal_ct_expand_batch");
    std::vector<uint64_t> ooffs(ncts + 1); std::vector<uint8_t> full(ncts * (113 + ctw * 8));
    if (batch(q, qbytes, coffs, ncts, e->N, e->p.primes, (uint32_t)e->L, full.data(), full.size(), ooffs.data(), 4)) return fail(e, PF_ERR_FORMAT, "query cts");
    std::vector<uint64_t> cts(ncts * ctw);
    uint64_t pid[4] = {0, 0, 0, 0};
    for (size_t c = 0; c < ncts; c++) {
        uint64_t n; int L, size, ntt;
        if (!pfo_ct_load(full.data() + ooffs[c], ooffs[c + 1] - ooffs[c], &n, &L, &size, &ntt, pid, cts.data() + c * ctw, ctw) || L != e->L || ntt) return fail(e, PF_ERR_FORMAT, "ct");
    }
    for (uint64_t i = 0; i < nq * nprobe; i++)
        if (idx[i] < 0 || (uint64_t)idx[i] >= e->nlist) return fail(e, PF_ERR_INVALID, "list id out of range");
    std::vector<int32_t> pq; std::vector<int64_t> pb; size_t nl = 0;
    for (uint64_t i = 0; i < nq; i++) {
        rpq[i] = 0; list_sizes[i] = 0;
        for (uint32_t p = 0; p < nprobe; p++) {
            const int64_t l = idx[i * nprobe + p];
            probed[i * nprobe + p] = (uint64_t)(e->offs[l + 1] - e->offs[l]); list_sizes[i] += probed[i * nprobe + p];
            for (int64_t o = e->offs[l]; o < e->offs[l + 1]; o++) { if (nl >= label_cap) return fail(e, PF_ERR_CAPACITY, "labels"); labels[nl++] = e->ids[o]; }
            for (size_t b = e->list_first_block[l]; b < e->list_first_block[l + 1]; b++) { pq.push_back((int32_t)i); pb.push_back((int64_t)b); rpq[i]++; }
        }
    }
    const size_t P = pq.size(), slot = pf_result_slot_size(e);
    if (P > max_results || P * slot > out_cap) return fail(e, PF_ERR_CAPACITY, "results");
    std::vector<const uint64_t *> keys;
    for (uint32_t r = 1; r < e->lay.R; r++) { auto it = e->keys.find(pfo_galois_elt_from_step(e->ctx, (int)r)); if (it == e->keys.end()) return fail(e, PF_ERR_STATE, "no key"); keys.push_back(it->second.data()); }
    std::vector<uint64_t> rot(nq * e->lay.K * ctw), res(P * 2 * e->Lr * e->N);
    double times[2];
    if (e->Lr < e->L) pfo_search_pairs_ms(e->ctx, &e->lay, nq, cts.data(), keys.data(), 0, P, pq.data(), pb.data(), e->diag.data(), e->norm.data(), rot.data(), e->Lr, res.data(), 8, times);
    else pfo_search_pairs(e->ctx, &e->lay, nq, cts.data(), keys.data(), 0, P, pq.data(), pb.data(), e->diag.data(), e->norm.data(), rot.data(), res.data(), 8, times);
    auto parms = sym<int (*)(uint64_t, const uint64_t *, uint32_t, uint64_t, uint64_t *)>("pf_parms_id");
    uint64_t opid[4]; if (e->Lr < e->L) parms(e->N, e->p.primes, (uint32_t)e->Lr, e->p.plain_modulus, opid); else memcpy(opid, pid, 32);
    memset(out, 0, P * slot);
    for (size_t r = 0; r < P; r++) { roffs[r] = r * slot + 15; pfo_ct_save(res.data() + r * 2 * e->Lr * e->N, e->N, e->Lr, 2, 0, opid, out + r * slot + 15); }
    roffs[P] = P * slot;
    if (st) { st->nresults = P; st->out_bytes = P * slot; st->useful_distances = nl; st->slot_distances = P * e->lay.C; }
    *ticket = 1; return PF_OK;
}
int pf_search_collect(pf_engine *, uint64_t) { return PF_OK; }
}
