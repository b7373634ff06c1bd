"""Reader / writer for the FAISS `IndexIVFPQ` file the reference server caches
(ref: src/server/server_lib.cpp:38-42 file name, :82 faiss::write_index, :91 faiss::read_index).

[EXT, UNVERIFIED] FAISS is not available in this container: the layout below is restated from the
published faiss/impl/index_write.cpp / index_read.cpp and could only be checked
against itself (round trip) — not against a file produced by FAISS.

What the GPU engine needs from the file: the coarse centroids (nested IndexFlatL2) and the inverted
lists (ids per list).  PQ codebooks and codes are carried through untouched so a file can be
re-written; the encrypted path works on the raw base vectors addressed by id, exactly as
Server::preciseSearch does (ref: src/server/server_lib.cpp:154-156).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


def _fourcc(s: str) -> int:
    return struct.unpack("<I", s.encode())[0]


@dataclass
class IVFPQFile:
    d: int
    ntotal: int
    nlist: int
    nprobe: int
    centroids: np.ndarray                      # [nlist][d] float32 (quantizer->reconstruct)
    list_ids: list                             # nlist arrays of int64
    list_codes: list                           # nlist arrays of uint8 [n][code_size]
    code_size: int = 32
    pq_M: int = 32
    pq_nbits: int = 8
    pq_centroids: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    by_residual: bool = True
    metric_type: int = 1                       # METRIC_L2
    is_trained: bool = True

    def csr(self, base_vectors: np.ndarray):
        """-> (list_offsets[nlist+1], ids, vectors in list order) for Engine.load_index"""
        sizes = np.array([len(x) for x in self.list_ids], dtype=np.int64)
        offsets = np.zeros(self.nlist + 1, dtype=np.int64)
        np.cumsum(sizes, out=offsets[1:])
        ids = np.concatenate(self.list_ids) if self.ntotal else np.zeros(0, np.int64)
        return offsets, ids.astype(np.int64), np.ascontiguousarray(base_vectors[ids], dtype=np.float32)


class _R:
    def __init__(self, b: bytes):
        self.b, self.o = b, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def arr(self, dtype, n):
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o).copy()
        self.o += a.nbytes
        return a

    def vec(self, dtype):
        return self.arr(dtype, self.take("Q"))


def _read_header(r: _R):
    d = r.take("i")
    ntotal = r.take("q")
    r.take("q")
    r.take("q")
    is_trained = bool(r.take("B"))
    metric = r.take("i")
    if metric > 1:
        r.take("f")
    return d, ntotal, is_trained, metric


def read_ivfpq(path: str) -> IVFPQFile:
    r = _R(open(path, "rb").read())
    h = r.take("I")
    if h != _fourcc("IwPQ"):
        raise ValueError("Loaded index is not of type IndexIVFPQ")  # ref: src/server/server_lib.cpp:92-95
    d, ntotal, is_trained, metric = _read_header(r)
    nlist, nprobe = r.take("Q"), r.take("Q")
    hq = r.take("I")
    if hq not in (_fourcc("IxF2"), _fourcc("IxFI")):
        raise ValueError("coarse quantizer is not an IndexFlat")
    dq, nq, _, _ = _read_header(r)
    cent = r.vec(np.float32)
    if dq != d or nq != nlist or cent.size != nlist * d:
        raise ValueError("quantizer shape does not match the IVF header")
    dm_type = r.take("B")
    r.vec(np.int64)
    if dm_type == 2:                                   # DirectMap::Hashtable
        n = r.take("Q")
        r.arr(np.int64, 2 * n)
    by_residual = bool(r.take("B"))
    code_size = r.take("Q")
    pq_d, pq_M, pq_nbits = r.take("Q"), r.take("Q"), r.take("Q")
    pq_cent = r.vec(np.float32)
    if r.take("I") != _fourcc("ilar"):
        raise ValueError("unsupported inverted-list container")
    il_nlist, il_cs = r.take("Q"), r.take("Q")
    if il_nlist != nlist or il_cs != code_size or pq_d != d:
        raise ValueError("inverted lists do not match the index header")
    kind = r.take("I")
    sizes = np.zeros(nlist, dtype=np.int64)
    if kind == _fourcc("full"):
        s = r.vec(np.uint64)
        if s.size != nlist:
            raise ValueError("bad list size table")
        sizes[:] = s
    elif kind == _fourcc("sprs"):
        s = r.vec(np.uint64).reshape(-1, 2)
        sizes[s[:, 0].astype(np.int64)] = s[:, 1]
    else:
        raise ValueError("unknown list size encoding")
    ids, codes = [], []
    for l in range(nlist):
        n = int(sizes[l])
        codes.append(r.arr(np.uint8, n * code_size).reshape(n, code_size))
        ids.append(r.arr(np.int64, n))
    if sum(len(x) for x in ids) != ntotal:
        raise ValueError("ntotal does not match the inverted lists")
    return IVFPQFile(d, ntotal, nlist, nprobe, cent.reshape(nlist, d), ids, codes, code_size, pq_M, pq_nbits, pq_cent,
                     by_residual, metric, is_trained)


def write_ivfpq(path: str, f: IVFPQFile):
    out = bytearray()

    def header(d, ntotal, trained, metric):
        return struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, int(trained), metric)

    out += struct.pack("<I", _fourcc("IwPQ")) + header(f.d, f.ntotal, f.is_trained, f.metric_type)
    out += struct.pack("<QQ", f.nlist, f.nprobe)
    cent = np.ascontiguousarray(f.centroids, dtype=np.float32)
    out += struct.pack("<I", _fourcc("IxF2")) + header(f.d, f.nlist, True, f.metric_type)
    out += struct.pack("<Q", cent.size) + cent.tobytes()
    out += struct.pack("<B", 0) + struct.pack("<Q", 0)                      # DirectMap::NoMap, empty array
    out += struct.pack("<BQ", int(f.by_residual), f.code_size)
    pqc = np.ascontiguousarray(f.pq_centroids, dtype=np.float32)
    out += struct.pack("<QQQ", f.d, f.pq_M, f.pq_nbits) + struct.pack("<Q", pqc.size) + pqc.tobytes()
    out += struct.pack("<I", _fourcc("ilar")) + struct.pack("<QQ", f.nlist, f.code_size)
    sizes = np.array([len(x) for x in f.list_ids], dtype=np.uint64)
    if np.count_nonzero(sizes) > f.nlist // 2:
        out += struct.pack("<I", _fourcc("full")) + struct.pack("<Q", f.nlist) + sizes.tobytes()
    else:
        nz = np.flatnonzero(sizes)
        pairs = np.stack([nz.astype(np.uint64), sizes[nz]], axis=1)
        out += struct.pack("<I", _fourcc("sprs")) + struct.pack("<Q", pairs.size) + pairs.tobytes()
    for l in range(f.nlist):
        if len(f.list_ids[l]):
            out += np.ascontiguousarray(f.list_codes[l], dtype=np.uint8).tobytes()
            out += np.ascontiguousarray(f.list_ids[l], dtype=np.int64).tobytes()
    open(path, "wb").write(bytes(out))
