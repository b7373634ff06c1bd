"""The PQ-ADC restatement of the FAISS fork's search_encrypted (oracle/pf_oracle.c: pfo_search_lists_pq,
pfo_pq_encode_residuals) against an independent float32 evaluation in numpy, and the .faiss file
layer carrying the quantizer.  [EXT]: the fork's source is absent, so this pins the oracle to the published
algorithm as restated, not to a FAISS build."""
import numpy as np

from tests.util import build_ivf, sift_like


def _case(seed, nb, d, nlist, nq, M):
    rng = np.random.default_rng(seed)
    base, query, cent = sift_like(rng, nb, d, nlist, nq)
    cent = (cent + rng.uniform(-0.5, 0.5, size=cent.shape)).astype(np.float32)
    offsets, ids, vecs = build_ivf(base, cent)
    list_of = np.repeat(np.arange(nlist), np.diff(offsets))
    dsub = d // M
    pick = rng.integers(0, len(vecs), size=(M, 256))
    resid = vecs - cent[list_of]
    pqc = np.stack([resid[pick[m], m * dsub:(m + 1) * dsub] for m in range(M)]).astype(np.float32)
    pqc += rng.normal(0, 0.25, size=pqc.shape).astype(np.float32)
    return query, cent, offsets, ids, vecs, list_of, pqc


def _table(r, pqc):
    """tab[m][j] = sum_k (r_mk - pq_mjk)^2 in float32, k ascending, multiply and add rounded separately"""
    M, ksub, dsub = pqc.shape
    tab = np.zeros((M, ksub), dtype=np.float32)
    rr = r.reshape(M, dsub)
    for k in range(dsub):
        t = (rr[:, None, k] - pqc[:, :, k]).astype(np.float32)
        tab = (tab + (t * t).astype(np.float32)).astype(np.float32)
    return tab


def test_pq_encode_and_adc_against_numpy_float32(oracle):
    d, M, nlist = 32, 8, 6
    query, cent, offsets, ids, vecs, list_of, pqc = _case(3, 400, d, nlist, 3, M)
    codes = oracle.pq_encode_residuals(vecs, offsets, cent, M, pqc)
    # encoder: nearest sub-centroid of the residual, first minimum wins
    for v in range(0, len(vecs), 7):
        r = (vecs[v] - cent[list_of[v]]).astype(np.float32)
        assert np.array_equal(codes[v], _table(r, pqc).argmin(1))
    idx = np.array([[0, 5, 2], [3, 3, 1], [4, 0, 5]], dtype=np.int64)        # a list probed twice is scored twice
    dist, labels, sizes = oracle.search_lists_pq(query, idx, cent, offsets, ids, M, pqc, codes)
    want_d, want_l = [], []
    for qi in range(3):
        for l in idx[qi]:
            r = (query[qi] - cent[l]).astype(np.float32)
            tab = _table(r, pqc)
            for o in range(int(offsets[l]), int(offsets[l + 1])):
                acc = np.float32(0)
                for m in range(M):
                    acc = np.float32(acc + tab[m, codes[o, m]])
                want_d.append(acc)
                want_l.append(ids[o])
    assert np.array_equal(labels, np.array(want_l)) and np.array_equal(dist.view(np.uint32), np.array(want_d, dtype=np.float32).view(np.uint32))
    assert np.array_equal(sizes, [sum(int(offsets[l + 1] - offsets[l]) for l in idx[qi]) for qi in range(3)])
    # and it is the squared L2 to the decoded vector (float64) up to float rounding
    dec = cent[list_of].astype(np.float64) + np.concatenate([pqc[m, codes[:, m]] for m in range(M)], axis=1)
    o = 0
    pos = {int(i): k for k, i in enumerate(ids)}
    for qi in range(3):
        n = int(sizes[qi])
        rows = [pos[int(i)] for i in labels[o:o + n]]
        assert np.allclose(dist[o:o + n], ((dec[rows] - query[qi]) ** 2).sum(1), rtol=2e-5, atol=1e-2)
        o += n


def test_faiss_file_carries_the_quantizer(oracle, tmp_path):
    from prefhetch_b200 import faiss_io
    d, M, nlist = 128, 32, 10           # the reference's shape: 32 sub-quantizers of 8 bits over 128 dimensions
    query, cent, offsets, ids, vecs, list_of, pqc = _case(4, 600, d, nlist, 2, M)
    codes = oracle.pq_encode_residuals(vecs, offsets, cent, M, pqc)
    lists = [ids[offsets[l]:offsets[l + 1]] for l in range(nlist)]
    f = faiss_io.IVFPQFile(d, len(ids), nlist, 20, cent, lists, [codes[offsets[l]:offsets[l + 1]] for l in range(nlist)],
                           code_size=M, pq_M=M, pq_nbits=8, pq_centroids=pqc.reshape(-1))
    faiss_io.write_ivfpq(str(tmp_path / "a.faiss"), f)
    g = faiss_io.read_ivfpq(str(tmp_path / "a.faiss"))
    assert (g.pq_M, g.pq_nbits, g.code_size) == (M, 8, M)
    assert np.array_equal(np.asarray(g.pq_centroids).reshape(M, 256, d // M), pqc)
    assert np.array_equal(np.concatenate(g.list_codes), codes)
    idx, _ = oracle.coarse_quantize(query, cent, 4)
    a = oracle.search_lists_pq(query, idx, cent, offsets, ids, M, pqc, codes)
    b = oracle.search_lists_pq(query, idx, g.centroids, offsets, np.concatenate(g.list_ids), g.pq_M, g.pq_centroids, np.concatenate(g.list_codes))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
