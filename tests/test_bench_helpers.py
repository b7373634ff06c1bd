"""CPU checks of bench.py's untimed bookkeeping (no GPU): the recall@10 computation against a hand-made
case, and the reference arm's light data set generator."""
import numpy as np

import bench


class _StubEngine:
    """stage 1 / plaintext stage 2 through the oracle (test infrastructure), same return layout as Engine"""

    def __init__(self, oracle, data, drop_label=None):
        self.o, self.d, self.drop = oracle, data, drop_label

    def coarse_quantize(self, x, nprobe):
        return self.o.coarse_quantize(x, self.d["centroids"], nprobe)[0]

    def coarseSearch(self, x, idx):
        dist, labels, sizes = self.o.search_lists_plain(x, idx, self.d["offsets"], self.d["ids"], self.d["vectors"])
        if self.drop is not None:                       # make one true neighbour look far away
            dist = dist.copy()
            dist[labels == self.drop] = 1e7
        return dist, labels, sizes


def test_recall_at_10_counts_hits(oracle):
    cfg = dict(nb=6000, d=32, nlist=16)
    data = bench.make_dataset(cfg, "cpu")
    full = bench.recall_at_10(_StubEngine(oracle, data), data, nprobe=16, dev="cpu", nq_r=8)
    assert full == 1.0                                   # every list probed: stage 2 sees the whole base set
    # brute-force top-10 of query 0, then hide its nearest neighbour from stage 2: exactly one miss in 80
    x = data["queries"][:1].astype(np.int64)
    d2 = ((data["vectors"].astype(np.int64) - x) ** 2).sum(1)
    key = d2 * (1 << 21) + data["ids"]
    nearest = int(np.sort(key)[0] % (1 << 21))
    miss = bench.recall_at_10(_StubEngine(oracle, data, drop_label=nearest), data, nprobe=16, dev="cpu", nq_r=8)
    hidden_in_others = sum(nearest in set((np.sort(((data["vectors"].astype(np.int64) - data["queries"][i:i + 1].astype(np.int64)) ** 2).sum(1)
                                                     * (1 << 21) + data["ids"])[:10] % (1 << 21)).tolist()) for i in range(1, 8))
    assert abs(miss - (1.0 - (1 + hidden_in_others) / 80.0)) < 1e-12
    few = bench.recall_at_10(_StubEngine(oracle, data), data, nprobe=1, dev="cpu", nq_r=8)
    assert 0.0 < few <= 1.0


def test_cpu_light_dataset_is_a_valid_ivf():
    cfg = dict(nb=5000, d=16, nlist=8)
    data = bench.make_dataset_cpu_light(cfg)
    off, ids, vec, cent = data["offsets"], data["ids"], data["vectors"], data["centroids"]
    assert off[0] == 0 and off[-1] == cfg["nb"] and np.all(np.diff(off) >= 0)
    assert sorted(ids.tolist()) == list(range(cfg["nb"]))
    assert vec.min() >= 0 and vec.max() <= 255 and np.array_equal(vec, np.rint(vec))
    # every vector sits in the list of its nearest centroid
    lab = ((vec[:, None, :] - cent[None, :, :]) ** 2).sum(-1).argmin(1)
    owner = np.repeat(np.arange(cfg["nlist"]), np.diff(off))
    assert (lab == owner).mean() > 0.999                 # float rounding may flip an exact tie
