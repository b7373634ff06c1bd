"""FAISS IndexIVFPQ file layer: round trip and error behaviour.  [EXT: cannot be
checked against a FAISS-written file here]"""
import numpy as np
import pytest

from prefhetch_b200 import faiss_io
from tests.util import build_ivf, sift_like


def _make(tmp_path, sparse=False):
    rng = np.random.default_rng(3)
    base, _, cent = sift_like(rng, 700, 128, 16, 1)
    offsets, ids, vecs = build_ivf(base, cent)
    list_ids = [ids[offsets[l]:offsets[l + 1]].copy() for l in range(16)]
    if sparse:
        for l in range(3, 16):
            list_ids[l] = np.zeros(0, np.int64)
    codes = [rng.integers(0, 256, size=(len(x), 32), dtype=np.uint8) for x in list_ids]
    f = faiss_io.IVFPQFile(128, sum(len(x) for x in list_ids), 16, 20, cent, list_ids, codes,
                           pq_centroids=rng.random(32 * 256 * 4, dtype=np.float32))
    p = tmp_path / "NBASE10000_PRECISE_DIMENSIONS_IVF256_PQ32_SUB_QUANTIZER_SIZE8.faiss"
    faiss_io.write_ivfpq(str(p), f)
    return f, p, base


@pytest.mark.parametrize("sparse", [False, True])
def test_roundtrip(tmp_path, sparse):
    f, p, base = _make(tmp_path, sparse)
    g = faiss_io.read_ivfpq(str(p))
    assert (g.d, g.ntotal, g.nlist, g.nprobe, g.code_size) == (f.d, f.ntotal, f.nlist, f.nprobe, f.code_size)
    assert np.array_equal(g.centroids, f.centroids) and np.array_equal(g.pq_centroids, f.pq_centroids)
    for a, b, ca, cb in zip(g.list_ids, f.list_ids, g.list_codes, f.list_codes):
        assert np.array_equal(a, b) and np.array_equal(ca, cb)
    offsets, ids, vecs = g.csr(base)
    assert offsets[-1] == g.ntotal and np.array_equal(vecs, base[ids])
    # byte-identical re-write
    p2 = tmp_path / "again.faiss"
    faiss_io.write_ivfpq(str(p2), g)
    assert p.read_bytes() == p2.read_bytes()


def test_wrong_type_and_truncation(tmp_path):
    f, p, _ = _make(tmp_path)
    raw = bytearray(p.read_bytes())
    raw[:4] = b"IxF2"
    (tmp_path / "flat.faiss").write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="not of type IndexIVFPQ"):   # ref: server_lib.cpp:92-95
        faiss_io.read_ivfpq(str(tmp_path / "flat.faiss"))
    (tmp_path / "cut.faiss").write_bytes(p.read_bytes()[:-100])
    with pytest.raises(Exception):
        faiss_io.read_ivfpq(str(tmp_path / "cut.faiss"))


def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & (2**64 - 1)
    return h


@pytest.mark.parametrize("sparse", [False, True])
def test_cpp_reader_matches_python(tmp_path, sparse):
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parents[1] / "prefhetch_b200" / "host" / "pf_faiss_check"
    if not exe.exists():
        import __graft_entry__
        __graft_entry__.build()
    f, p, base = _make(tmp_path, sparse)
    out = subprocess.run([str(exe), str(p)], capture_output=True, text=True, check=True).stdout.split()
    offsets, ids, _ = f.csr(base)
    assert [int(x) for x in out[:5]] == [f.d, f.ntotal, f.nlist, f.nprobe, f.code_size]
    assert int(out[5], 16) == _fnv(np.ascontiguousarray(f.centroids, np.float32).tobytes())
    assert int(out[6], 16) == _fnv(offsets.tobytes()) and int(out[7], 16) == _fnv(ids.tobytes())
    # the product quantizer and the codes (list order) for pf_load_pq
    assert [int(out[8]), int(out[9])] == [f.pq_M, f.pq_nbits]
    assert int(out[10], 16) == _fnv(np.ascontiguousarray(f.pq_centroids, np.float32).tobytes())
    assert int(out[11], 16) == _fnv(np.concatenate([np.asarray(c, np.uint8).reshape(-1) for c in f.list_codes]).tobytes())
    bad = tmp_path / "bad.faiss"
    bad.write_bytes(b"IxF2" + p.read_bytes()[4:])
    r = subprocess.run([str(exe), str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "not of type IndexIVFPQ" in r.stderr


def _variant(tmp_path, name, **kw):
    """the base file with one field of the IVFPQFile changed, written, read back by both readers"""
    import dataclasses
    f, p, base = _make(tmp_path)
    g = dataclasses.replace(f, **kw)
    q = tmp_path / name
    faiss_io.write_ivfpq(str(q), g)
    return g, q, base


@pytest.mark.parametrize("case", ["array_map", "hash_map", "ivfpqr"])
def test_accepts_what_the_reference_cast_accepts(tmp_path, case):
    """direct maps of every kind and IndexIVFPQR (a subclass of IndexIVFPQ: the reference's dynamic_cast takes it,
    ref: src/server/server_lib.cpp:92-95) load; a re-write keeps their bytes; the C++ reader agrees"""
    import subprocess
    from pathlib import Path
    rng = np.random.default_rng(9)
    kw = {"array_map": dict(direct_map_type=1, direct_map_array=rng.integers(0, 1 << 40, size=700).astype(np.int64)),
          "hash_map": dict(direct_map_type=2, direct_map_pairs=rng.integers(0, 1 << 40, size=(37, 2)).astype(np.int64)),
          # refine_pq {d, M, nbits, centroids}, refine_codes (vector<uint8>), k_factor (float)
          "ivfpqr": dict(refine=np.array([128, 8, 8], np.uint64).tobytes() + np.array([5], np.uint64).tobytes()
                         + np.arange(5, dtype=np.float32).tobytes() + np.array([16], np.uint64).tobytes() + bytes(16)
                         + np.array([4.0], np.float32).tobytes())}[case]
    g, q, base = _variant(tmp_path, case + ".faiss", **kw)
    h = faiss_io.read_ivfpq(str(q))
    assert (h.d, h.ntotal, h.nlist) == (g.d, g.ntotal, g.nlist) and np.array_equal(h.centroids, g.centroids)
    assert all(np.array_equal(a, b) for a, b in zip(h.list_ids, g.list_ids))
    assert h.direct_map_type == g.direct_map_type and np.array_equal(h.direct_map_array, g.direct_map_array)
    assert np.array_equal(h.direct_map_pairs, g.direct_map_pairs) and h.refine == g.refine
    q2 = tmp_path / "again.faiss"
    faiss_io.write_ivfpq(str(q2), h)
    assert q.read_bytes() == q2.read_bytes()
    exe = Path(__file__).resolve().parents[1] / "prefhetch_b200" / "host" / "pf_faiss_check"
    out = subprocess.run([str(exe), str(q)], capture_output=True, text=True, check=True).stdout.split()
    offsets, ids, _ = g.csr(base)
    assert [int(x) for x in out[:3]] == [g.d, g.ntotal, g.nlist] and int(out[7], 16) == _fnv(ids.tobytes())


@pytest.mark.parametrize("case,msg", [("legacy", "legacy"), ("il00", "no inverted lists"), ("ilod", "unsupported inverted-list"),
                                      ("badmap", "direct-map"), ("trailing", "trailing bytes"), ("quantizer", "IndexFlat"),
                                      ("sprs_oob", "list size table"), ("cut_lists", "truncated"), ("ntotal", "ntotal")])
def test_rejects_with_a_reason(tmp_path, case, msg):
    """every refusal names its reason in both readers (Python ValueError, C++ std::runtime_error -> exit 1)"""
    import subprocess
    from pathlib import Path
    f, p, _ = _make(tmp_path, sparse=(case == "sprs_oob"))
    raw = bytearray(p.read_bytes())
    ilar = bytes(raw).index(b"ilar")
    if case == "legacy":
        raw[:4] = b"IvPQ"
    elif case == "il00":
        raw[ilar:ilar + 4] = b"il00"
    elif case == "ilod":
        raw[ilar:ilar + 4] = b"ilod"
    elif case == "badmap":
        dm = 4 + 33 + 16 + 4 + 33 + 8 + 16 * 128 * 4        # fourcc, header, nlist/nprobe, IxF2, header, count, centroids
        assert raw[dm] == 0
        raw[dm] = 7
    elif case == "trailing":
        raw += b"\\0" * 8
    elif case == "quantizer":
        iq = 4 + 33 + 16
        raw[iq:iq + 4] = b"IxPQ"
    elif case == "sprs_oob":
        sp = bytes(raw).index(b"sprs") + 4 + 8
        raw[sp:sp + 8] = (99).to_bytes(8, "little")           # list number beyond nlist
    elif case == "cut_lists":
        raw = raw[:ilar + 200]
    elif case == "ntotal":
        raw[8:16] = (f.ntotal + 1).to_bytes(8, "little")
    q = tmp_path / "bad.faiss"
    q.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match=msg):
        faiss_io.read_ivfpq(str(q))
    exe = Path(__file__).resolve().parents[1] / "prefhetch_b200" / "host" / "pf_faiss_check"
    r = subprocess.run([str(exe), str(q)], capture_output=True, text=True)
    assert r.returncode == 1 and msg in r.stderr
