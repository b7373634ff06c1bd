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
    # what the reference's dynamic_cast<IndexIVFPQ*> also accepts and a re-write must keep:
    direct_map_type: int = 0                   # DirectMap::NoMap / Array (1) / Hashtable (2)
    direct_map_array: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    direct_map_pairs: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int64))
    refine: bytes | None = None                # IndexIVFPQR ("IwQR"): refine_pq + refine_codes + k_factor, kept verbatim

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
        size = struct.calcsize("<" + fmt)
        if self.o + size > len(self.b):
            raise ValueError("truncated index file")
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += size
        return v if len(v) > 1 else v[0]

    def arr(self, dtype, n):
        nbytes = int(n) * np.dtype(dtype).itemsize
        if n < 0 or self.o + nbytes > len(self.b):
            raise ValueError("truncated index file")
        a = np.frombuffer(self.b, dtype=dtype, count=int(n), offset=self.o).copy()
        self.o += nbytes
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
    # IndexIVFPQR ("IwQR") derives from IndexIVFPQ: the reference's dynamic_cast accepts it (ref:
    # src/server/server_lib.cpp:92-95); its refine section follows the inverted lists and is carried verbatim.
    # "IvPQ" / "IvQR" are the pre-2018 layouts (lists stored per list with their own headers): refused by name.
    if h in (_fourcc("IvPQ"), _fourcc("IvQR")):
        raise ValueError("legacy IndexIVFPQ file layout (IvPQ / IvQR): re-save it with a current FAISS")
    if h not in (_fourcc("IwPQ"), _fourcc("IwQR")):
        raise ValueError("Loaded index is not of type IndexIVFPQ")  # ref: src/server/server_lib.cpp:92-95
    is_pqr = h == _fourcc("IwQR")
    d, ntotal, is_trained, metric = _read_header(r)
    nlist, nprobe = r.take("Q"), r.take("Q")
    if d <= 0 or ntotal < 0 or nlist == 0 or nlist > (1 << 32):
        raise ValueError("implausible index header")
    hq = r.take("I")
    if hq not in (_fourcc("IxF2"), _fourcc("IxFI"), _fourcc("IxFl")):
        raise ValueError("coarse quantizer is not an IndexFlat")
    dq, nq, _, _ = _read_header(r)
    cent = r.vec(np.float32)
    if dq != d or nq != nlist or cent.size != nlist * d:
        raise ValueError("quantizer shape does not match the IVF header")
    dm_type = r.take("B")
    if dm_type > 2:
        raise ValueError("unknown direct-map type")
    dm_array = r.vec(np.int64)
    dm_pairs = np.zeros((0, 2), np.int64)
    if dm_type == 2:                                   # DirectMap::Hashtable: vector of (id, list/offset) pairs
        n = r.take("Q")
        dm_pairs = r.arr(np.int64, 2 * n).reshape(-1, 2)
    by_residual = bool(r.take("B"))
    code_size = r.take("Q")
    pq_d, pq_M, pq_nbits = r.take("Q"), r.take("Q"), r.take("Q")
    pq_cent = r.vec(np.float32)
    il = r.take("I")
    if il == _fourcc("il00"):
        raise ValueError("index file holds no inverted lists (il00)")
    if il != _fourcc("ilar"):
        raise ValueError("unsupported inverted-list container (only in-memory ArrayInvertedLists, 'ilar')")
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
        s = r.vec(np.uint64)
        if s.size % 2:
            raise ValueError("bad list size table")
        s = s.reshape(-1, 2)
        if s.size and int(s[:, 0].max()) >= nlist:
            raise ValueError("bad list size table")
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
    refine = None
    if is_pqr:
        refine = bytes(r.b[r.o:])
        if len(refine) < 24 + 8 + 8 + 4:
            raise ValueError("truncated index file")
    elif r.o != len(r.b):
        raise ValueError("trailing bytes after the inverted lists")
    return IVFPQFile(d, ntotal, nlist, nprobe, cent.reshape(nlist, d), ids, codes, code_size, pq_M, pq_nbits, pq_cent,
                     by_residual, metric, is_trained, dm_type, dm_array, dm_pairs, refine)


def write_ivfpq(path: str, f: IVFPQFile):
    out = bytearray()

    def header(d, ntotal, trained, metric):
        return struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, int(trained), metric)

    out += struct.pack("<I", _fourcc("IwQR" if f.refine is not None else "IwPQ")) + header(f.d, f.ntotal, f.is_trained, f.metric_type)
    out += struct.pack("<QQ", f.nlist, f.nprobe)
    cent = np.ascontiguousarray(f.centroids, dtype=np.float32)
    out += struct.pack("<I", _fourcc("IxF2")) + header(f.d, f.nlist, True, f.metric_type)
    out += struct.pack("<Q", cent.size) + cent.tobytes()
    dma = np.ascontiguousarray(f.direct_map_array, dtype=np.int64)
    out += struct.pack("<B", f.direct_map_type) + struct.pack("<Q", dma.size) + dma.tobytes()
    if f.direct_map_type == 2:
        dmp = np.ascontiguousarray(f.direct_map_pairs, dtype=np.int64)
        out += struct.pack("<Q", dmp.shape[0]) + dmp.tobytes()
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
    if f.refine is not None:
        out += f.refine
    open(path, "wb").write(bytes(out))
