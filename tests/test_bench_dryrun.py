"""bench.py's control flow on the CPU (PF_BENCH_DRYRUN=1: gloo, stub engine — tests/bench_stub.py): every rank
takes the same path through the collectives at 1, 2 and 4 ranks, on the default run (weak headline + strong record,
2 x N/2 grid, rank 0's solo 1-GPU reference), on explicit grids, with the shared-memory response buffer; and a
failure or a hang injected into an optional stage on one rank still yields the headline line and exit code 0."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


_ports_lock, _ports_used = __import__("threading").Lock(), set()


def _port():
    """a free rendezvous port, never handed out twice in this process (scenarios start concurrently)"""
    with _ports_lock:
        while True:
            with socket.socket() as s:
                s.bind(("127.0.0.1", 0))
                p = s.getsockname()[1]
            if p not in _ports_used:
                _ports_used.add(p)
                return p


def _run(n, extra, env_extra=None, timeout=900):
    env = dict(os.environ, PF_BENCH_DRYRUN="1", PF_BENCH_DRYRUN_NB="12000", PF_BENCH_CLOCKS="off", OMP_NUM_THREADS="1")
    env.update(env_extra or {})
    base = [sys.executable]
    if n > 1:
        base += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                 "--master-port", str(_port())]
    cmd = base + [str(ROOT / "bench.py"), "--gpus", str(n), "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-parity"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=str(ROOT))
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    return r, (json.loads(lines[-1]) if lines else None)


# Every scenario is one bench.py job (1 - 4 processes that mostly wait on each other); they are independent, so the
# module runs them four at a time up front and the tests below only look at the results.
SCENARIOS = {
    "single": (1, [], None),
    "default2": (2, [], None),
    "default4": (4, [], None),
    "grid1x4": (4, ["--config", "sift1m_nlist4096_nprobe64", "--grid", "1x4"], None),
    "grid4x1": (4, ["--config", "sift1m_nlist4096_nprobe64", "--grid", "4x1"], None),
    "fail:1": (2, ["--no-strong"], {"PF_BENCH_DRYRUN_INJECT": "fail:1", "PF_BENCH_STAGE_LIMIT_S": "8"}),
    "hang:1": (2, ["--no-strong"], {"PF_BENCH_DRYRUN_INJECT": "hang:1", "PF_BENCH_STAGE_LIMIT_S": "8"}),
    "fail:0": (2, ["--no-strong"], {"PF_BENCH_DRYRUN_INJECT": "fail:0", "PF_BENCH_STAGE_LIMIT_S": "8"}),
    "configs4": (2, ["--no-strong"], {"PF_BENCH_EXTRAS_MIN_GPUS": "2"}),
    "short_allowance": (2, [], {"PF_BENCH_EXTRAS_MIN_GPUS": "2", "PF_BENCH_RUN_LIMIT_S": "200"}),
    "allowance_fires": (2, ["--no-strong"], {"PF_BENCH_DRYRUN_INJECT": "hang:1", "PF_BENCH_STAGE_LIMIT_S": "600", "PF_BENCH_RUN_LIMIT_S": "60"}),
    "private_response": (2, ["--no-strong"], {"PF_BENCH_PRIVATE_RESPONSE": "1"}),
}


@pytest.fixture(scope="module")
def runs():
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=4) as pool:
        futures = {name: pool.submit(_run, n, extra, env, 600) for name, (n, extra, env) in SCENARIOS.items()}
        results = {}
        for name, f in futures.items():
            try:
                results[name] = f.result()
            except Exception as ex:     # noqa: BLE001 — reported by the test that asks for this scenario
                results[name] = ex

    def get(name):
        r = results[name]
        if isinstance(r, Exception):
            raise r
        return r
    return get


def test_single_rank_line(runs):
    r, line = runs("single")
    assert r.returncode == 0 and line is not None, r.stderr[-2000:]
    assert line["n_gpus"] == 1 and line["e2e"]["value"] > 0 and line.get("aborted_stage") is None
    assert line["recall_at_10"] > 0.9 and 0.0 < line["recall_overlapping_mixture"]["recall_at_10"] <= 1.0
    assert line["config"]["parallelism"] == "single" and line["roofline"]["launches_per_step"] == 1.0


@pytest.mark.parametrize("n", [2, 4])
def test_default_multi_rank_run(runs, n):
    """weak headline + strong record (grid 2 x N/2) + gather verification + one shared response buffer"""
    r, line = runs(f"default{n}")
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["n_gpus"] == n and line["scaling"] == "weak" and line.get("aborted_stage") is None
    assert line["gather_verified"]["ranks"] == n - 1 and "mismatch_ranks" not in line["gather_verified"]
    assert line["e2e"]["value"] > 0 and "one host buffer per node" in line["e2e"]["response"]
    s = line["strong"]
    assert s["n_gpus"] == n and s["value"] > 0 and s["value_1gpu"] > 0 and s["gather_verified"]["ranks"] == n - 1
    assert f"2 list shards x {n // 2} query groups" in s["config"]["parallelism"]
    assert s["e2e"]["value"] > 0 and s["e2e_1gpu"]["value"] > 0


@pytest.mark.parametrize("grid", ["1x4", "4x1"])
def test_explicit_grids(runs, grid):
    r, line = runs(f"grid{grid}")
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["scaling"] == "strong" and line["gather_verified"]["ranks"] == 3 and line["e2e"]["value"] > 0


@pytest.mark.parametrize("inject", ["fail:1", "hang:1", "fail:0"])
def test_optional_stage_failure_keeps_the_headline(runs, inject):
    """an exception or a hang inside e2e on one rank: the line published after the timed region is emitted with
    `aborted_stage`, every rank exits 0"""
    r, line = runs(inject)
    assert r.returncode == 0, r.stderr[-3000:]
    assert line is not None and line["value"] > 0 and line["n_gpus"] == 2
    assert line["aborted_stage"] is not None and "e2e" in line["aborted_stage"]["stage"]


def test_configs4_record_stage(runs):
    """the 8-GPU default run records BASELINE configs[4] as a last optional stage (here triggered at 2 ranks)"""
    r, line = runs("configs4")
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    c4 = line["configs4"]
    assert c4["value"] > 0 and c4["config"]["workload"] == "synth10m_nlist16384" and c4["config"]["queries_per_step"] == 256
    assert c4["gather_verified"]["ranks"] == 1 and line.get("aborted_stage") is None


def test_extras_are_skipped_when_the_run_allowance_is_short(runs):
    """the driver kills a run after its per-N limit: with little of the allowance left the strong record and the
    configs[4] record are skipped (rank 0 decides, every rank agrees) and the headline is emitted as usual"""
    r, line = runs("short_allowance")
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["value"] > 0 and line.get("aborted_stage") is None
    assert "skipped" in line["strong"] and "skipped" in line["configs4"]


def test_run_allowance_fires_the_guard(runs):
    """the overall allowance ends while an optional stage is still running: the published headline goes out"""
    r, line = runs("allowance_fires")
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["value"] > 0 and "allowance" in line["aborted_stage"]["why"]


def test_private_response_fallback(runs):
    """no room in /dev/shm for one node buffer: every rank keeps its share in private pinned memory, completion flags
    stay shared; the e2e number is still produced and says so"""
    r, line = runs("private_response")
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["e2e"]["value"] > 0 and "per-rank pinned buffers" in line["e2e"]["response"] and line.get("aborted_stage") is None
