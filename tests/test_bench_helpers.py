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


def test_build_line_has_the_contract_keys():
    """the JSON line is assembled from run_workload's record: every key of the bench contract is there, with an
    optional stage that failed recorded instead of raising, and it serialises"""
    import argparse
    import json
    args = argparse.Namespace(steps=20, warmup=5)
    cfg_name = "sift1m_nlist1024_nprobe16"
    cfg = dict(bench.CONFIGS[cfg_name])
    rec = {"value": 4.0e8, "ms_per_step": 2.5, "steps": 20, "warmup": 5, "useful_per_step": 1.0e6, "slot_distances_per_s": 4.5e8,
           "result_cts_per_step": 1100.0, "gpu_launches": 700, "roofline": {"bound": "hbm", "achieved": 4500.0, "peak": 6548.2, "unit": "GB/s",
                                                                             "frac": 0.69, "traffic": None},
           "rotate_roofline": {"bound": "fp64+imad pipes", "rotations_per_step": 960, "ms_per_step": 1.0, "fp64_ops_per_step": 4.1e9,
                               "imad_wide_per_step": 9.4e8, "lanes_per_clk_per_sm": 64, "sms": 148},
           "phases_ms_per_step": {"mac": 0.83}, "queries_per_s": 25000.0, "timed_window": (0.0, 1.0), "gather_verified": None,
           "info": {}, "L": 4, "Lr": 1, "nprobe": 16, "db_gib_per_rank": 4.5,
           "e2e": {"error": "RuntimeError('x')"}, "parity": {"error": "boom", "parity_checked": 0}}
    clocks = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}
    line = bench.build_line(args, cfg_name, cfg, 1, False, (1, 1), rec, clocks, {"pinned": False}, None)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "roofline", "e2e", "cpu_baseline", "gpu_launches", "parity_checked"):
        assert key in line, key
    assert line["config"]["workload"] == cfg_name and "model" not in line["config"]
    assert line["scaling"] == "weak" and line["parity_checked"] == 0 and line["vs_baseline"] is None
    assert 0.0 < line["rotate_roofline"]["frac"] < 1.0
    json.dumps(line)
    strong = {"scaling": "strong", "efficiency": 0.8}
    line8 = bench.build_line(args, cfg_name, cfg, 8, True, (8, 1), rec, clocks, {}, strong)
    assert line8["scaling"] == "weak" and line8["strong"]["efficiency"] == 0.8 and line8["config"]["nlist"] == 8192
    line_s = bench.build_line(args, "sift1m_nlist4096_nprobe64", dict(bench.CONFIGS["sift1m_nlist4096_nprobe64"]), 8, False, (2, 4), rec, clocks, {}, None)
    assert line_s["scaling"] == "strong" and "2 list shards x 4 query groups" in line_s["config"]["parallelism"]


def test_reference_arm_names_the_same_config_as_the_gpu_arm():
    """both arms build `config` with the same function and the same arguments -> same_config"""
    cfg_name = "sift1m_nlist1024_nprobe16"
    cfg = dict(bench.CONFIGS[cfg_name])
    a = bench.bench_config(cfg_name, cfg, 4, 1, cfg["nq"])                       # --impl reference
    b = bench.bench_config(cfg_name, cfg, 4, 1, cfg["nq"], 1, False, 16, (1, 1))  # ours, N = 1
    assert a == b
    # N = 4 default run: the GPU arm weak-scales (4 list shards, nprobe 64); run_reference builds the same object
    ours = bench.bench_config(cfg_name, cfg, 4, 1, cfg["nq"], 4, True, 64, (4, 1))
    ref = bench.bench_config(cfg_name, cfg, 4, 1, cfg["nq"], 4, True, min(cfg["nprobe"] * 4, cfg["nlist"]), (4, 1))
    assert ours == ref and ours["nprobe"] == 64 and ours["nb"] == 4 * cfg["nb"]
    assert bench.cpu_sample_queries(64, 16) == 32 and bench.cpu_sample_queries(64, 1) == 8 and bench.cpu_sample_queries(16, 64) == 16


def test_overlapping_mixture_has_recall_below_one(oracle):
    """bench.py reports recall@10 on SURVEY's well-separated mixture (1.0) AND on an overlapping one, where probing
    a few lists misses true neighbours; the reference definition (|GT100 ∩ top10| / 10) is never below the standard one"""
    cfg = dict(nb=20000, d=32, nlist=64)
    hard = bench.make_dataset(cfg, "cpu", seed=777, sigma=48.0, spread=128.0, lloyd=2, offset=64.0)
    off = hard["offsets"]
    assert off[-1] == cfg["nb"] and hard["vectors"].min() >= 0 and hard["vectors"].max() <= 255
    rm = bench.recall_metrics(_StubEngine(oracle, hard), hard, nprobe=2, dev="cpu", nq_r=32)
    assert 0.3 < rm["recall_at_10"] < 0.98 and rm["reference_recall_10"] >= rm["recall_at_10"]
    full = bench.recall_metrics(_StubEngine(oracle, hard), hard, nprobe=64, dev="cpu", nq_r=8)
    assert full["recall_at_10"] == 1.0


def test_bench_calls_bind_to_the_real_engine():
    """bench.py's multi-GPU flow is dry-run on the CPU against tests/bench_stub.py; this keeps the stub honest: every
    `eng.<method>(...)` call in bench.py binds to the signature of the REAL prefhetch_b200.Engine method (count and
    keyword names), every attribute bench.py reads exists on the real class or is set in its __init__, and every stub
    method has the real one's arity."""
    import ast
    import inspect
    from pathlib import Path
    from prefhetch_b200.engine import Engine
    from tests import bench_stub
    root = Path(__file__).resolve().parent.parent
    tree = ast.parse((root / "bench.py").read_text())
    init_src = inspect.getsource(Engine.__init__) + inspect.getsource(Engine.set_list_sizes)
    calls, attrs = [], set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "eng":
            attrs.add(node.attr)
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) \
                and node.func.value.id == "eng":
            calls.append((node.lineno, node.func.attr, len(node.args), [k.arg for k in node.keywords if k.arg]))
    assert len(calls) > 40
    for lineno, name, npos, kws in calls:
        assert hasattr(Engine, name), f"bench.py:{lineno}: Engine has no method {name}"
        inspect.signature(getattr(Engine, name)).bind(None, *([0] * npos), **{k: 0 for k in kws})
    for a in attrs:
        assert hasattr(Engine, a) or f"self.{a}" in init_src, f"bench.py reads eng.{a}, which the real Engine does not have"
    stub = bench_stub.Engine
    for name, fn in inspect.getmembers(stub, inspect.isfunction):
        if name.startswith("_"):
            continue
        assert hasattr(Engine, name), f"stub method {name} does not exist on the real Engine"
        def arity(f):
            ps = [q for q in inspect.signature(f).parameters.values() if q.name != "self" and q.kind == q.POSITIONAL_OR_KEYWORD]
            return sum(q.default is q.empty for q in ps), len(ps)
        req_s, max_s = arity(fn)
        req_r, max_r = arity(getattr(Engine, name))
        assert req_s == req_r and max_s <= max_r, f"stub {name}: takes {req_s}..{max_s} arguments, the real method {req_r}..{max_r}"
    # the constructor call of bench.py binds too
    ctor = [n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr == "Engine"]
    assert ctor
    for n in ctor:
        inspect.signature(Engine.__init__).bind(None, *([0] * len(n.args)), **{k.arg: 0 for k in n.keywords if k.arg})


def test_reference_functions_timed_uses_the_reference_itself():
    """the configs[0] reference arm times the reference's own compiled functions (oracle/_ref) when they exist"""
    import bench
    from oracle import pf_ref as R
    rng = np.random.default_rng(0)
    data = {"vectors": rng.integers(0, 256, size=(3000, 128)).astype(np.float32), "queries": rng.integers(0, 256, size=(8, 128)).astype(np.float32),
            "centroids": rng.uniform(0, 200, size=(50, 128)).astype(np.float32)}
    r = bench.reference_functions_timed(data)
    if R.build():
        assert r["kind"] == "reference" and r["cores"] == 1 and r["exact_distances_per_s"] > 1e4 and r["centroid_distances_per_s"] > 1e4
    else:
        assert "unavailable" in r
    data["vectors"] = data["vectors"][:, :64]
    assert "unavailable" in bench.reference_functions_timed(data)
