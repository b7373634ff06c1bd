"""Python host mirror of the reference `Server` class over the C ABI.

Method names and argument meaning follow the reference (ref: include/server/server_lib.h:25-49,
src/server/server_lib.cpp:101-167), with run-time shapes instead of compile-time std::array, and
RuntimeError-derived exceptions where the reference throws std::runtime_error.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import PfIndexInfo, PfParams, PfSearchStats

# SEAL util/globals.cpp: CoeffModulus::BFVDefault(N) for 128-bit security
_BFV_DEFAULT = {
    4096: [0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001],
    8192: [0x7FFFFFD8001, 0x7FFFFFC8001, 0xFFFFFFFC001, 0xFFFFFF6C001, 0xFFFFFEBC001],
    16384: [0xFFFFFFFD8001, 0xFFFFFFFA0001, 0xFFFFFFF00001, 0x1FFFFFFF68001, 0x1FFFFFFF50001,
            0x1FFFFFFEE8001, 0x1FFFFFFEA0001, 0x1FFFFFFE88001, 0x1FFFFFFE48001],
}
# SEAL PlainModulus::Batching(N, bits)
_BATCHING = {(4096, 20): 1032193, (4096, 24): 16760833, (8192, 24): 16760833, (8192, 27): 133857281,
             (16384, 24): 16580609, (16384, 27): 133857281}


def bfv_default_primes(n: int):
    return list(_BFV_DEFAULT[n])


def parms_id(poly_degree: int, primes, plain_modulus: int):
    """SEAL parms_id (4 x uint64) of a BFV parameter set; needs no GPU (pf_parms_id)"""
    lib = _capi.load()
    pr = (C.c_uint64 * len(primes))(*primes)
    out = (C.c_uint64 * 4)()
    if lib.pf_parms_id(poly_degree, pr, len(primes), plain_modulus, out):
        raise ValueError("pf_parms_id: invalid arguments")
    return tuple(int(x) for x in out)


def seal_stream_inflate(blob) -> bytes:
    """one SEAL stream -> its compr_mode none form (zlib and zstd streams inflated); needs no GPU"""
    lib = _capi.load()
    b = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
    need, used = C.c_size_t(), C.c_size_t()
    rc = lib.pf_seal_stream_inflate(b.ctypes.data_as(C.c_void_p), b.size, None, 0, C.byref(need), C.byref(used))
    if rc != _capi.PF_ERR_CAPACITY:
        raise PfError(rc, "malformed SEAL stream")
    out = np.empty(need.value, dtype=np.uint8)
    rc = lib.pf_seal_stream_inflate(b.ctypes.data_as(C.c_void_p), b.size, out.ctypes.data_as(C.c_void_p), out.size,
                                    C.byref(need), C.byref(used))
    if rc:
        raise PfError(rc, "malformed SEAL stream")
    return out.tobytes()


def seal_ct_expand(blob, poly_degree: int, data_primes) -> bytes:
    """a SEAL ciphertext stream (seeded and / or zlib / zstd) -> the equivalent full compr_mode none stream; needs no GPU
    (pf_seal_ct_expand)"""
    lib = _capi.load()
    b = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
    pr = (C.c_uint64 * len(data_primes))(*data_primes)
    need, used = C.c_size_t(), C.c_size_t()
    rc = lib.pf_seal_ct_expand(b.ctypes.data_as(C.c_void_p), b.size, poly_degree, pr, len(data_primes), None, 0,
                               C.byref(need), C.byref(used))
    if rc != _capi.PF_ERR_CAPACITY:
        raise PfError(rc, "malformed or unsupported SEAL ciphertext stream")
    out = np.empty(need.value, dtype=np.uint8)
    rc = lib.pf_seal_ct_expand(b.ctypes.data_as(C.c_void_p), b.size, poly_degree, pr, len(data_primes),
                               out.ctypes.data_as(C.c_void_p), out.size, C.byref(need), C.byref(used))
    if rc:
        raise PfError(rc, "malformed or unsupported SEAL ciphertext stream")
    return out.tobytes()


def seal_ct_expand_batch(blob, offsets, poly_degree: int, data_primes, threads: int = 0):
    """every stream of a request (seeded and / or compressed) -> (full compr_mode none streams back to back, offsets);
    the host slow path of the search calls, multi-threaded; needs no GPU (pf_seal_ct_expand_batch)"""
    lib = _capi.load()
    b = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    ncts = len(offs) - 1
    pr = (C.c_uint64 * len(data_primes))(*data_primes)
    out_offs = np.zeros(ncts + 1, dtype=np.uint64)
    rc = lib.pf_seal_ct_expand_batch(b.ctypes.data_as(C.c_void_p), b.size, _ptr(offs, U64P), ncts, poly_degree, pr, len(data_primes),
                                     None, 0, _ptr(out_offs, U64P), threads)
    if rc not in (_capi.PF_OK, _capi.PF_ERR_CAPACITY):
        raise PfError(rc, "malformed or unsupported SEAL ciphertext stream in the batch")
    out = np.empty(max(1, int(out_offs[-1])), dtype=np.uint8)
    rc = lib.pf_seal_ct_expand_batch(b.ctypes.data_as(C.c_void_p), b.size, _ptr(offs, U64P), ncts, poly_degree, pr, len(data_primes),
                                     out.ctypes.data_as(C.c_void_p), out.size, _ptr(out_offs, U64P), threads)
    if rc:
        raise PfError(rc, "malformed or unsupported SEAL ciphertext stream in the batch")
    return out[:int(out_offs[-1])].tobytes(), out_offs


def seal_galois_keys_expand(blob, poly_degree: int, key_primes) -> bytes:
    """a SEAL GaloisKeys stream (Serializable<GaloisKeys>: seeded key ciphertexts; and / or zlib / zstd) -> the
    equivalent full compr_mode none stream over the k key primes; needs no GPU (pf_seal_galois_keys_expand)"""
    lib = _capi.load()
    b = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
    pr = (C.c_uint64 * len(key_primes))(*key_primes)
    need = C.c_size_t()
    rc = lib.pf_seal_galois_keys_expand(b.ctypes.data_as(C.c_void_p), b.size, poly_degree, pr, len(key_primes), None, 0, C.byref(need))
    if rc != _capi.PF_ERR_CAPACITY:
        raise PfError(rc, "malformed or unsupported SEAL GaloisKeys stream")
    out = np.empty(need.value, dtype=np.uint8)
    rc = lib.pf_seal_galois_keys_expand(b.ctypes.data_as(C.c_void_p), b.size, poly_degree, pr, len(key_primes),
                                        out.ctypes.data_as(C.c_void_p), out.size, C.byref(need))
    if rc:
        raise PfError(rc, "malformed or unsupported SEAL GaloisKeys stream")
    return out.tobytes()


def batching_plain_modulus(n: int, bits: int) -> int:
    return _BATCHING[(n, bits)]


class PfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[pf {code}] {msg}")
        self.code = code


@dataclass
class SearchResult:
    """Encrypted stage-2 response: SEAL-serialized result ciphertexts plus the plaintext envelope."""
    blob: np.ndarray                # uint8, result ciphertexts in aligned slots
    result_offsets: np.ndarray      # [nresults+1]: start of every SEAL stream; last = bytes used
    ct_bytes: int                   # length of every result stream
    results_per_query: np.ndarray   # [nq]
    labels: np.ndarray              # ids of the owned probed lists, packed per query
    list_sizes: np.ndarray          # [nq]
    probed_sizes: np.ndarray        # [nq][nprobe]
    stats: dict

    def result(self, r: int) -> bytes:
        o = int(self.result_offsets[r])
        return self.blob[o:o + self.ct_bytes].tobytes()


class PendingSearch:
    """A search in flight (pf_search_submit); collect() = pf_search_collect."""

    def __init__(self, eng, ticket: int, result: SearchResult, keep):
        self.eng, self.ticket, self.result, self._keep = eng, ticket, result, keep

    def collect(self) -> SearchResult:
        if self.ticket is not None:
            self.eng._ck(self.eng.lib.pf_search_collect(self.eng.h, self.ticket))
            self.ticket, self._keep = None, None
        return self.result


def _ptr(a: np.ndarray, typ):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ)


U64P, I64P, F32P, U8P, I32P = (C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_float),
                               C.POINTER(C.c_uint8), C.POINTER(C.c_int32))


class Engine:
    """One engine per GPU (one process per GPU).  Mirrors `Server` (ref: include/server/server_lib.h)."""

    def __init__(self, dim: int, poly_degree: int = 8192, primes=None, plain_modulus: int | None = None,
                 query_cts: int = 1, partial_g: int = 8, device: int = 0, rank: int = 0, world: int = 1,
                 result_limbs: int = 0):
        self.lib = _capi.load()
        primes = list(primes) if primes is not None else bfv_default_primes(poly_degree)
        if plain_modulus is None:
            plain_modulus = batching_plain_modulus(poly_degree, 24)
        p = PfParams()
        p.struct_size = C.sizeof(PfParams)
        p.device, p.poly_degree, p.num_primes, p.dim = device, poly_degree, len(primes), dim
        for i, q in enumerate(primes):
            p.primes[i] = q
        p.plain_modulus, p.query_cts, p.partial_g, p.rank, p.world = plain_modulus, query_cts, partial_g, rank, world
        p.result_limbs = result_limbs
        self.h = C.c_void_p()
        rc = self.lib.pf_engine_create(C.byref(p), C.byref(self.h))
        if rc:
            raise PfError(rc, (self.lib.pf_last_error(None) or b"").decode())
        self.n, self.primes, self.t, self.dim = poly_degree, primes, plain_modulus, dim
        self.k, self.L = len(primes), len(primes) - 1
        self.m, self.g = query_cts, partial_g
        self.ctw = 2 * self.L * self.n
        self.ct_bytes = self.lib.pf_ct_serialized_size(self.h)
        self.slot_bytes = self.lib.pf_result_slot_size(self.h)
        self.result_bytes = self.lib.pf_result_serialized_size(self.h)
        self.Lr = result_limbs or self.L
        self.device, self.rank, self.world = device, rank, world

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.pf_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc: int):
        if rc:
            raise PfError(rc, (self.lib.pf_last_error(self.h) or b"").decode())

    # ---- index -------------------------------------------------------------------------------
    def load_index(self, centroids, list_offsets, ids, vectors):
        c = np.ascontiguousarray(centroids, dtype=np.float32)
        lo = np.ascontiguousarray(list_offsets, dtype=np.int64)
        i = np.ascontiguousarray(ids, dtype=np.int64)
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        assert c.shape[1] == self.dim and v.shape[1] == self.dim and len(lo) == c.shape[0] + 1
        self._ck(self.lib.pf_load_index(self.h, c.shape[0], _ptr(c, F32P), _ptr(lo, I64P), _ptr(i, I64P),
                                        _ptr(v, F32P)))
        self.nlist = c.shape[0]
        return self.index_info()

    def load_index_from_faiss(self, path: str, base_vectors):
        """Server::init_index's cached-file branch (ref: src/server/server_lib.cpp:88-100): centroids and
        inverted lists come from the .faiss file, raw vectors from the base set addressed by id."""
        from . import faiss_io
        f = faiss_io.read_ivfpq(path)
        if f.d != self.dim:
            raise PfError(1, f"index dimension {f.d} does not match the engine ({self.dim})")
        offsets, ids, vecs = f.csr(np.asarray(base_vectors))
        info = self.load_index(f.centroids, offsets, ids, vecs)
        # the file's product quantizer (what the reference's search_encrypted scores with today), when it holds one
        if f.pq_nbits == 8 and f.pq_M and f.pq_centroids.size == 256 * f.d and len(f.list_codes) == f.nlist:
            codes = np.concatenate([np.asarray(c, dtype=np.uint8).reshape(-1, f.pq_M) for c in f.list_codes]) if f.ntotal else \
                np.zeros((0, f.pq_M), np.uint8)
            self.load_pq(f.pq_M, f.pq_nbits, f.pq_centroids, codes)
        return info

    def index_info(self) -> dict:
        o = PfIndexInfo()
        self._ck(self.lib.pf_get_index_info(self.h, C.byref(o)))
        return {f: getattr(o, f) for f, _ in PfIndexInfo._fields_}

    def retrieve_centroids(self) -> np.ndarray:
        """ref: Server::retrieve_centroids (src/server/server_lib.cpp:101-109)"""
        out = np.zeros((self.nlist, self.dim), dtype=np.float32)
        self._ck(self.lib.pf_retrieve_centroids(self.h, _ptr(out, F32P), out.size))
        return out

    # ---- stage 1 -----------------------------------------------------------------------------
    def coarse_quantize(self, x, nprobe: int, return_dist: bool = False):
        """ref: sort_nearest_centroids + first NPROBE (src/client/client_lib.cpp:50-81, :93-103)"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        idx = np.zeros((x.shape[0], nprobe), dtype=np.int64)
        dist = np.zeros((x.shape[0], nprobe), dtype=np.float32)
        self._ck(self.lib.pf_coarse_quantize(self.h, x.shape[0], _ptr(x, F32P), nprobe, _ptr(idx, I64P),
                                             _ptr(dist, F32P)))
        return (idx, dist) if return_dist else idx

    # ---- stage 2 plaintext -------------------------------------------------------------------
    def coarseSearch(self, precise_query, nearest_centroid_idx):
        """ref: Server::coarseSearch (src/server/server_lib.cpp:111-138): returns
        (coarse_distance_scores, coarse_distance_indexes, list_sizes_per_query)."""
        x = np.ascontiguousarray(precise_query, dtype=np.float32)
        idx = np.ascontiguousarray(nearest_centroid_idx, dtype=np.int64)
        nq, nprobe = idx.shape
        sizes = np.zeros(nq, dtype=np.uint64)
        total = C.c_uint64()
        rc = self.lib.pf_search_lists_plain(self.h, nq, _ptr(x, F32P), _ptr(idx, I64P), nprobe, None, None, 0,
                                            _ptr(sizes, U64P), C.byref(total))
        if rc not in (_capi.PF_OK, _capi.PF_ERR_CAPACITY):
            self._ck(rc)
        dist = np.zeros(max(total.value, 1), dtype=np.float32)
        labels = np.zeros(max(total.value, 1), dtype=np.int64)
        self._ck(self.lib.pf_search_lists_plain(self.h, nq, _ptr(x, F32P), _ptr(idx, I64P), nprobe, _ptr(dist, F32P),
                                                _ptr(labels, I64P), total.value, _ptr(sizes, U64P), C.byref(total)))
        return dist[:total.value], labels[:total.value], sizes.astype(np.int64)

    def load_pq(self, pq_M: int, pq_nbits: int, pq_centroids, codes):
        """product quantizer of the loaded index (ref: the IndexIVFPQ built at src/server/server_lib.cpp:34-36): the
        `pq_M`, `pq_nbits`, `pq_centroids` of a .faiss file (faiss_io.IVFPQFile) and its list codes back to back"""
        pqc = np.ascontiguousarray(pq_centroids, dtype=np.float32).reshape(-1)
        cd = np.ascontiguousarray(codes, dtype=np.uint8)
        if pqc.size != 256 * self.dim or cd.ndim != 2 or cd.shape[1] != pq_M:
            raise PfError(_capi.PF_ERR_INVALID, "pq_centroids must hold 256 * dim floats and codes must be [ntotal][pq_M]")
        self._ck(self.lib.pf_load_pq(self.h, pq_M, pq_nbits, _ptr(pqc, F32P), cd.ctypes.data_as(C.POINTER(C.c_uint8))))

    def coarseSearchPQ(self, precise_query, nearest_centroid_idx):
        """Server::coarseSearch with the distance the reference's FAISS fork computes today (PQ-ADC over every code
        of the given lists; ref: src/server/server_lib.cpp:111-138, the search_encrypted call at :126-130): same
        packing as coarseSearch."""
        x = np.ascontiguousarray(precise_query, dtype=np.float32)
        idx = np.ascontiguousarray(nearest_centroid_idx, dtype=np.int64)
        nq, nprobe = idx.shape
        sizes = np.zeros(nq, dtype=np.uint64)
        total = C.c_uint64()
        rc = self.lib.pf_search_lists_pq(self.h, nq, _ptr(x, F32P), _ptr(idx, I64P), nprobe, None, None, 0,
                                         _ptr(sizes, U64P), C.byref(total))
        if rc not in (_capi.PF_OK, _capi.PF_ERR_CAPACITY):
            self._ck(rc)
        dist = np.zeros(max(total.value, 1), dtype=np.float32)
        labels = np.zeros(max(total.value, 1), dtype=np.int64)
        self._ck(self.lib.pf_search_lists_pq(self.h, nq, _ptr(x, F32P), _ptr(idx, I64P), nprobe, _ptr(dist, F32P),
                                             _ptr(labels, I64P), total.value, _ptr(sizes, U64P), C.byref(total)))
        return dist[:total.value], labels[:total.value], sizes.astype(np.int64)

    def ct_expand_seeded_device(self, stream) -> np.ndarray:
        """pf_seal_ct_expand_device: an uncompressed blake2xb-seeded ciphertext stream -> words [2][L][n], c1 drawn on
        the device (what the search calls do for all-seeded requests)"""
        b = np.frombuffer(bytes(stream), dtype=np.uint8)
        out = np.zeros((2, self.L, self.n), dtype=np.uint64)
        self._ck(self.lib.pf_seal_ct_expand_device(self.h, b.ctypes.data_as(C.c_void_p), b.size, _ptr(out, U64P), out.size))
        return out

    def preciseSearch(self, precise_query, nearest_coarse_vector_idx) -> np.ndarray:
        """ref: Server::preciseSearch (src/server/server_lib.cpp:140-167)"""
        x = np.ascontiguousarray(precise_query, dtype=np.float32)
        ids = np.ascontiguousarray(nearest_coarse_vector_idx, dtype=np.int64)
        out = np.zeros(ids.shape, dtype=np.float32)
        self._ck(self.lib.pf_precise_search(self.h, ids.shape[0], _ptr(x, F32P), _ptr(ids, I64P), ids.shape[1],
                                            _ptr(out, F32P)))
        return out

    # ---- Galois keys -------------------------------------------------------------------------
    def galois_elt(self, step: int) -> int:
        return self.lib.pf_galois_elt_from_step(self.h, step)

    def set_galois_key(self, elt: int, key_words: np.ndarray):
        kw = np.ascontiguousarray(key_words, dtype=np.uint64)
        assert kw.size == self.L * 2 * self.k * self.n
        self._ck(self.lib.pf_set_galois_key(self.h, elt, _ptr(kw, U64P)))

    def load_galois_keys(self, blob: bytes):
        b = np.frombuffer(blob, dtype=np.uint8)
        self._ck(self.lib.pf_load_galois_keys(self.h, _ptr(np.ascontiguousarray(b), U8P), len(b)))

    # ---- stage 2 encrypted -------------------------------------------------------------------
    def coarseSearchEncrypted(self, query_blob, ct_offsets, nearest_centroid_idx, out: np.ndarray | None = None
                              ) -> SearchResult:
        """Encrypted variant of Server::coarseSearch: SEAL-serialized query ciphertexts in, SEAL-serialized
        result ciphertexts out (additive to ref: src/server/controllers/Query.cc:29-63)."""
        return self.submitSearchEncrypted(query_blob, ct_offsets, nearest_centroid_idx, out).collect()

    def submitSearchEncrypted(self, query_blob, ct_offsets, nearest_centroid_idx, out: np.ndarray | None = None
                              ) -> "PendingSearch":
        """pf_search_submit: enqueue the search and return; `.collect()` waits for the result ciphertexts.
        Two searches may be in flight (upload + compute of one overlap the download of the other).  The
        query blob and `out` must not be touched until collect() returns."""
        qb = query_blob if isinstance(query_blob, np.ndarray) else np.frombuffer(query_blob, dtype=np.uint8)
        offs = np.ascontiguousarray(ct_offsets, dtype=np.uint64)
        idx = np.ascontiguousarray(nearest_centroid_idx, dtype=np.int64)
        nq, nprobe = idx.shape
        if not hasattr(self, "_C"):
            self._C = self.index_info()["C"]
        # worst case: every probed list owned here
        max_results = int(self._max_results(idx))
        if out is None:
            out = np.zeros(max(1, max_results) * self.slot_bytes, dtype=np.uint8)
        label_cap = max(1, max_results * self._C)
        # response arrays are reused across calls (fresh 9 MB label arrays cost more page faults than the
        # copy itself): a ring of 5 sets (one more than the searches that can be in flight), so the views of a
        # result stay valid while the next four searches are submitted
        key = (nq, nprobe)
        ring = getattr(self, "_enc_bufs", None)
        if ring is None or ring["key"] != key or len(ring["sets"][0]["roff"]) < max_results + 1 \
                or len(ring["sets"][0]["labels"]) < label_cap:
            ring = {"key": key, "next": 0, "sets": [
                {"roff": np.empty(max_results + 1 + 64, dtype=np.uint64), "rpq": np.empty(nq, dtype=np.uint64),
                 "labels": np.empty(label_cap + label_cap // 8, dtype=np.int64), "sizes": np.empty(nq, dtype=np.uint64),
                 "psz": np.empty((nq, nprobe), dtype=np.uint64)} for _ in range(5)]}
            self._enc_bufs = ring
        bufs = ring["sets"][ring["next"]]
        roff, rpq, labels, sizes, psz = bufs["roff"], bufs["rpq"], bufs["labels"], bufs["sizes"], bufs["psz"]
        label_cap = len(labels)
        st = PfSearchStats()
        ticket = C.c_uint64()
        self._ck(self.lib.pf_search_submit(
            self.h, nq, qb.ctypes.data_as(C.c_void_p), qb.size, _ptr(offs, U64P), _ptr(idx, I64P), nprobe,
            out.ctypes.data_as(C.c_void_p), out.size, _ptr(roff, U64P), max_results, _ptr(rpq, U64P),
            _ptr(labels, I64P), label_cap, _ptr(sizes, U64P), _ptr(psz, U64P), C.byref(st), C.byref(ticket)))
        ring["next"] = (ring["next"] + 1) % 5          # a refused submit does not consume a set
        nres = st.nresults
        res = SearchResult(out, roff[:nres + 1], self.result_bytes, rpq.astype(np.int64), labels[:int(sizes.sum())],
                           sizes.astype(np.int64), psz.astype(np.int64),
                           {f: getattr(st, f) for f, _ in PfSearchStats._fields_})
        return PendingSearch(self, ticket.value, res, (qb, offs, idx))

    def set_search_groups(self, groups: int):
        """pf_search_set_groups: 1 = whole-batch kernels (pipelined calls), 0 = default (4 query groups)"""
        self._ck(self.lib.pf_search_set_groups(self.h, groups))

    def host_register(self, arr: np.ndarray):
        self._ck(self.lib.pf_host_register(self.h, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self._ck(self.lib.pf_host_unregister(self.h, arr.ctypes.data_as(C.c_void_p)))

    def device_checksum(self, dptr: int, nwords: int, cuda_stream: int = 0) -> int:
        out = C.c_uint64()
        self._ck(self.lib.pf_device_checksum(self.h, C.c_void_p(dptr), nwords, C.byref(out), C.c_void_p(cuda_stream)))
        return out.value

    def _max_results(self, idx: np.ndarray) -> int:
        if not hasattr(self, "_blocks_per_list"):
            raise PfError(_capi.PF_ERR_STATE, "set_list_sizes() not called")
        return int(self._blocks_per_list[idx.reshape(-1)].sum())

    def set_list_sizes(self, list_offsets):
        """host-side copy of the list lengths for sizing response buffers"""
        lo = np.asarray(list_offsets, dtype=np.int64)
        C_ = self.index_info()["C"]
        sizes = lo[1:] - lo[:-1]
        owned = (np.arange(len(sizes)) % self.world) == self.rank
        self._blocks_per_list = np.where(owned, (sizes + C_ - 1) // C_, 0)

    def search_device(self, d_query_ptr: int, nq: int, idx, d_out_ptr: int, cap_results: int):
        """Device-resident step (asynchronous on the engine stream).  Pointers are raw device addresses."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        rpq = np.zeros(nq, dtype=np.uint64)
        st = PfSearchStats()
        self._ck(self.lib.pf_search_device(self.h, nq, C.c_void_p(d_query_ptr), _ptr(idx, I64P), idx.shape[1],
                                           C.c_void_p(d_out_ptr), cap_results, _ptr(rpq, U64P), C.byref(st)))
        return rpq.astype(np.int64), {f: getattr(st, f) for f, _ in PfSearchStats._fields_}

    # ---- peer-memory gather buffers (multi-GPU) ------------------------------------------------
    def ipc_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte handle) of a buffer other processes can map with ipc_open"""
        ptr = C.c_void_p()
        h = (C.c_uint8 * 64)()
        self._ck(self.lib.pf_ipc_alloc(self.h, nbytes, C.byref(ptr), h))
        return ptr.value, bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        h = (C.c_uint8 * 64)(*handle)
        self._ck(self.lib.pf_ipc_open(self.h, h, C.byref(ptr)))
        return ptr.value

    def copy_async(self, dst: int, src: int, nbytes: int, cuda_stream: int = 0):
        self._ck(self.lib.pf_copy_async(self.h, C.c_void_p(dst), C.c_void_p(src), nbytes, C.c_void_p(cuda_stream)))

    def flag_write(self, ptr: int, value: int, cuda_stream: int = 0):
        self._ck(self.lib.pf_flag_write(self.h, C.c_void_p(ptr), value, C.c_void_p(cuda_stream)))

    def flag_wait(self, ptr: int, value: int, cuda_stream: int = 0):
        self._ck(self.lib.pf_flag_wait(self.h, C.c_void_p(ptr), value, C.c_void_p(cuda_stream)))

    def ipc_close(self, ptr: int):
        self._ck(self.lib.pf_ipc_close(self.h, C.c_void_p(ptr)))

    def ipc_free(self, ptr: int):
        self._ck(self.lib.pf_ipc_free(self.h, C.c_void_p(ptr)))

    # ---- stream / timing ---------------------------------------------------------------------
    def stream(self) -> int:
        return self.lib.pf_engine_stream(self.h) or 0

    def set_stream(self, cuda_stream: int):
        self._ck(self.lib.pf_engine_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.lib.pf_engine_synchronize(self.h))

    def timing_enable(self, on: bool = True):
        self._ck(self.lib.pf_timing_enable(self.h, int(on)))

    def timing_read(self, reset: bool = True) -> dict:
        ms = (C.c_float * _capi.PF_T_COUNT)()
        ln = (C.c_uint64 * _capi.PF_T_COUNT)()
        self._ck(self.lib.pf_timing_read(self.h, ms, ln, int(reset)))
        return {name: {"ms": ms[i], "launches": ln[i]} for name, i in _capi.PHASES.items()}

    def launch_count(self) -> int:
        return self.lib.pf_launch_count(self.h)

    # ---- primitives (parity entry points) ----------------------------------------------------
    def ntt_forward(self, polys: np.ndarray, limbs) -> np.ndarray:
        a = np.ascontiguousarray(polys, dtype=np.uint64).copy().reshape(-1, self.n)
        lm = np.ascontiguousarray(limbs, dtype=np.int32)
        self._ck(self.lib.pf_ntt_forward(self.h, _ptr(a, U64P), a.shape[0], _ptr(lm, I32P)))
        return a

    def ntt_inverse(self, polys: np.ndarray, limbs) -> np.ndarray:
        a = np.ascontiguousarray(polys, dtype=np.uint64).copy().reshape(-1, self.n)
        lm = np.ascontiguousarray(limbs, dtype=np.int32)
        self._ck(self.lib.pf_ntt_inverse(self.h, _ptr(a, U64P), a.shape[0], _ptr(lm, I32P)))
        return a

    def ct_pt_mac(self, cts: np.ndarray, pts: np.ndarray, addend: np.ndarray | None = None) -> np.ndarray:
        c = np.ascontiguousarray(cts, dtype=np.uint64)
        p = np.ascontiguousarray(pts, dtype=np.uint64)
        out = np.zeros((2, self.L, self.n), dtype=np.uint64)
        ad = None if addend is None else _ptr(np.ascontiguousarray(addend, dtype=np.uint64), U64P)
        self._ck(self.lib.pf_ct_pt_mac(self.h, _ptr(c, U64P), _ptr(p, U64P), c.shape[0], ad, _ptr(out, U64P)))
        return out

    def ct_add(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        out = np.zeros((2, self.L, self.n), dtype=np.uint64)
        self._ck(self.lib.pf_ct_add(self.h, _ptr(np.ascontiguousarray(a, dtype=np.uint64), U64P),
                                    _ptr(np.ascontiguousarray(b, dtype=np.uint64), U64P), _ptr(out, U64P)))
        return out

    def ct_to_ntt(self, cts: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(cts, dtype=np.uint64).copy()
        self._ck(self.lib.pf_ct_to_ntt(self.h, _ptr(a, U64P), a.size // self.ctw))
        return a

    def ct_from_ntt(self, cts: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(cts, dtype=np.uint64).copy()
        self._ck(self.lib.pf_ct_from_ntt(self.h, _ptr(a, U64P), a.size // self.ctw))
        return a

    def rotate_rows(self, ct: np.ndarray, step: int) -> np.ndarray:
        out = np.zeros((2, self.L, self.n), dtype=np.uint64)
        self._ck(self.lib.pf_rotate_rows(self.h, _ptr(np.ascontiguousarray(ct, dtype=np.uint64), U64P), step,
                                         _ptr(out, U64P)))
        return out

    def rotate_query_set(self, cts: np.ndarray, chain: bool = False) -> np.ndarray:
        info = self.index_info()
        out = np.zeros((info["K"], 2, self.L, self.n), dtype=np.uint64)
        self._ck(self.lib.pf_rotate_query_set(self.h, _ptr(np.ascontiguousarray(cts, dtype=np.uint64), U64P),
                                              int(chain), _ptr(out, U64P)))
        return out

    def batch_encode(self, values: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        assert v.size == self.n
        out = np.zeros(self.n, dtype=np.uint64)
        self._ck(self.lib.pf_batch_encode(self.h, _ptr(v, U64P), _ptr(out, U64P)))
        return out

    def encode_block(self, xs: np.ndarray):
        x = np.ascontiguousarray(xs, dtype=np.int32).reshape(-1, self.dim)
        info = self.index_info()
        diag = np.zeros((info["K"], self.L, self.n), dtype=np.uint64)
        norm = np.zeros((self.L, self.n), dtype=np.uint64)
        self._ck(self.lib.pf_encode_block(self.h, _ptr(x, I32P), x.shape[0], _ptr(diag, U64P), _ptr(norm, U64P)))
        return diag, norm

    def ct_serialize(self, ct: np.ndarray, is_ntt: bool = False) -> bytes:
        out = np.zeros(self.ct_bytes, dtype=np.uint8)
        w = C.c_size_t()
        self._ck(self.lib.pf_ct_serialize(self.h, _ptr(np.ascontiguousarray(ct, dtype=np.uint64), U64P), int(is_ntt),
                                          _ptr(out, U8P), out.size, C.byref(w)))
        return out[:w.value].tobytes()

    def ct_deserialize(self, blob) -> tuple[np.ndarray, bool]:
        """-> (ct [2][limbs][n], is_ntt); limbs is L for queries, result_limbs for results"""
        b = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
        ct = np.zeros(2 * self.L * self.n, dtype=np.uint64)
        ntt, limbs, used = C.c_int32(), C.c_int32(), C.c_size_t()
        self._ck(self.lib.pf_ct_deserialize(self.h, _ptr(b, U8P), b.size, _ptr(ct, U64P), ct.size, C.byref(limbs),
                                            C.byref(ntt), C.byref(used)))
        return ct[:2 * limbs.value * self.n].reshape(2, limbs.value, self.n), bool(ntt.value)

    def set_result_parms_id(self, pid):
        arr = (C.c_uint64 * 4)(*pid)
        self._ck(self.lib.pf_set_result_parms_id(self.h, arr))
