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
    bad = tmp_path / "bad.faiss"
    bad.write_bytes(b"IxF2" + p.read_bytes()[4:])
    r = subprocess.run([str(exe), str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "not of type IndexIVFPQ" in r.stderr
