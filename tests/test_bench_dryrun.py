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


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


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


def test_single_rank_line():
    r, line = _run(1, [])
    assert r.returncode == 0 and line is not None, r.stderr[-2000:]
    assert line["n_gpus"] == 1 and line["e2e"]["value"] > 0 and line.get("aborted_stage") is None
    assert line["recall_at_10"] > 0.9 and 0.0 < line["recall_overlapping_mixture"]["recall_at_10"] <= 1.0
    assert line["config"]["parallelism"] == "single" and line["roofline"]["launches_per_step"] == 1.0


@pytest.mark.parametrize("n", [2, 4])
def test_default_multi_rank_run(n):
    """weak headline + strong record (grid 2 x N/2) + gather verification + one shared response buffer"""
    r, line = _run(n, [])
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["n_gpus"] == n and line["scaling"] == "weak" and line.get("aborted_stage") is None
    assert line["gather_verified"]["ranks"] == n - 1 and "mismatch_ranks" not in line["gather_verified"]
    assert line["e2e"]["value"] > 0 and "one host buffer per node" in line["e2e"]["response"]
    s = line["strong"]
    assert s["n_gpus"] == n and s["value"] > 0 and s["value_1gpu"] > 0 and s["gather_verified"]["ranks"] == n - 1
    assert f"2 list shards x {n // 2} query groups" in s["config"]["parallelism"]
    assert s["e2e"]["value"] > 0 and s["e2e_1gpu"]["value"] > 0


@pytest.mark.parametrize("grid", ["1x4", "4x1"])
def test_explicit_grids(grid):
    r, line = _run(4, ["--config", "sift1m_nlist4096_nprobe64", "--grid", grid])
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["scaling"] == "strong" and line["gather_verified"]["ranks"] == 3 and line["e2e"]["value"] > 0


@pytest.mark.parametrize("inject", ["fail:1", "hang:1", "fail:0"])
def test_optional_stage_failure_keeps_the_headline(inject):
    """an exception or a hang inside e2e on one rank: the line published after the timed region is emitted with
    `aborted_stage`, every rank exits 0"""
    r, line = _run(2, ["--no-strong"], {"PF_BENCH_DRYRUN_INJECT": inject, "PF_BENCH_STAGE_LIMIT_S": "8"}, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert line is not None and line["value"] > 0 and line["n_gpus"] == 2
    assert line["aborted_stage"] is not None and "e2e" in line["aborted_stage"]["stage"]


def test_configs4_record_stage():
    """the 8-GPU default run records BASELINE configs[4] as a last optional stage (here triggered at 2 ranks)"""
    r, line = _run(2, ["--no-strong"], {"PF_BENCH_EXTRAS_MIN_GPUS": "2"})
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    c4 = line["configs4"]
    assert c4["value"] > 0 and c4["config"]["workload"] == "synth10m_nlist16384" and c4["config"]["queries_per_step"] == 256
    assert c4["gather_verified"]["ranks"] == 1 and line.get("aborted_stage") is None


def test_extras_are_skipped_when_the_run_allowance_is_short():
    """the driver kills a run after its per-N limit: with little of the allowance left the strong record and the
    configs[4] record are skipped (rank 0 decides, every rank agrees) and the headline is emitted as usual"""
    r, line = _run(2, [], {"PF_BENCH_EXTRAS_MIN_GPUS": "2", "PF_BENCH_RUN_LIMIT_S": "200"})
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["value"] > 0 and line.get("aborted_stage") is None
    assert "skipped" in line["strong"] and "skipped" in line["configs4"]


def test_run_allowance_fires_the_guard():
    """the overall allowance ends while an optional stage is still running: the published headline goes out"""
    r, line = _run(2, ["--no-strong"], {"PF_BENCH_DRYRUN_INJECT": "hang:1", "PF_BENCH_STAGE_LIMIT_S": "600", "PF_BENCH_RUN_LIMIT_S": "35"},
                   timeout=300)
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["value"] > 0 and "allowance" in line["aborted_stage"]["why"]


def test_private_response_fallback():
    """no room in /dev/shm for one node buffer: every rank keeps its share in private pinned memory, completion flags
    stay shared; the e2e number is still produced and says so"""
    r, line = _run(2, ["--no-strong"], {"PF_BENCH_PRIVATE_RESPONSE": "1"})
    assert r.returncode == 0 and line is not None, r.stderr[-3000:]
    assert line["e2e"]["value"] > 0 and "per-rank pinned buffers" in line["e2e"]["response"] and line.get("aborted_stage") is None
