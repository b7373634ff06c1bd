"""world_size-2 checks of the multi-GPU host logic on CPU (gloo): list sharding by rank, the
result-count exchange and the gather-to-rank-0 of per-shard result ciphertexts (bench.py's
gather_results pattern) — with synthetic payloads standing in for ciphertexts (no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def shard_plan(idx, list_sizes, C, rank, world):
    """host-side mirror of plan_pairs() in pf_engine.cu: result ciphertexts of the lists rank owns"""
    plan = []
    for qi in range(idx.shape[0]):
        for l in idx[qi]:
            if l % world != rank:
                continue
            for b in range((int(list_sizes[l]) + C - 1) // C):
                plan.append((qi, int(l), b))
    return plan


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    nlist, nq, nprobe, C, words = 32, 6, 8, 64, 16
    list_sizes = rng.integers(0, 200, size=nlist)
    idx = np.stack([rng.choice(nlist, size=nprobe, replace=False) for _ in range(nq)])
    plan = shard_plan(idx, list_sizes, C, rank, world)
    # payload of result r = a deterministic function of (query, list, block): stands in for the ciphertext words
    mine = torch.tensor([[qi * 1_000_000 + l * 1000 + b] * words for qi, l, b in plan], dtype=torch.int64).reshape(-1, words)
    cnt = torch.tensor([len(plan)], dtype=torch.int64)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    gathered = {0: mine} if rank == 0 else None
    if rank == 0:
        for r in range(1, world):
            c = int(cnts[r].item())
            buf = torch.empty((c, words), dtype=torch.int64)
            if c:
                dist.recv(buf, src=r)
            gathered[r] = buf
    elif len(plan):
        dist.send(mine, dst=0)
    dist.barrier()
    if rank == 0:
        # the union over ranks must be exactly the single-rank plan, every (query, list, block) once
        full = shard_plan(idx, list_sizes, C, 0, 1)
        got = sorted(int(row[0]) for r in range(world) for row in gathered[r])
        want = sorted(qi * 1_000_000 + l * 1000 + b for qi, l, b in full)
        ok = got == want and sum(int(c.item()) for c in cnts) == len(full)
        with open(out_path, "w") as f:
            f.write("ok" if ok else f"mismatch {len(got)} {len(want)}")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_shard_and_gather_gloo(tmp_path, world):
    port = _free_port()
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, port, str(out)), nprocs=world, join=True)
    assert out.read_text() == "ok"


def test_shards_partition_the_lists():
    rng = np.random.default_rng(1)
    list_sizes = rng.integers(0, 3000, size=100)
    idx = np.stack([rng.choice(100, size=16, replace=False) for _ in range(5)])
    full = set(shard_plan(idx, list_sizes, 1024, 0, 1))
    for world in (2, 4, 8):
        parts = [set(shard_plan(idx, list_sizes, 1024, r, world)) for r in range(world)]
        assert set().union(*parts) == full
        assert sum(len(p) for p in parts) == len(full)


# ---- round 2: the rank grid (list shards x query groups) and the one-buffer node response ----------------
def test_grid_covers_every_query_list_pair_once():
    """bench.py's grid: rank r serves list shard r % L and query group r // L; every (query, probed list)
    pair is computed by exactly one rank, for every grid of 1, 2, 4 and 8 ranks"""
    import bench
    rng = np.random.default_rng(3)
    nq, nlist, nprobe = 64, 256, 16
    idx = np.stack([rng.choice(nlist, size=nprobe, replace=False) for _ in range(nq)])
    for world in (1, 2, 4, 8):
        grids = [(L, world // L) for L in (1, 2, 4, 8) if world % L == 0 and L <= world]
        assert bench.default_strong_grid(world) in grids
        for Lw, Qw in grids:
            assert bench.parse_grid(f"{Lw}x{Qw}", world) == (Lw, Qw)
            seen = np.zeros((nq, nlist), dtype=np.int32)
            for rank in range(world):
                lr, qg = rank % Lw, rank // Lw
                lo, hi = bench.split_queries(nq, Qw, qg)
                for qi in range(lo, hi):
                    for l in idx[qi]:
                        if l % Lw == lr:
                            seen[qi, l] += 1
            want = np.zeros_like(seen)
            for qi in range(nq):
                want[qi, idx[qi]] = 1
            assert np.array_equal(seen, want), (Lw, Qw)
    with pytest.raises(SystemExit):
        bench.parse_grid("3x2", 8)
    assert bench.parse_grid(None, 4) == (4, 1)


def _response_worker(rank, world, port, name, out_path):
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shares = [1000 + 64 * r for r in range(world)]
    qbytes = 5000
    if rank != 0:
        dist.barrier()                       # rank 0 creates the segment first
    nr = bench.NodeResponse(rank, world, name, qbytes, shares)
    if rank == 0:
        nr.flags[:] = 0
        nr.query[:] = np.arange(qbytes, dtype=np.uint32).astype(np.uint8)
        dist.barrier()
    dist.barrier()
    ok = nr.share.size == shares[rank] and np.array_equal(nr.query, np.arange(qbytes, dtype=np.uint32).astype(np.uint8))
    for step in range(3):                    # every rank writes its share of the response of `step`, then marks it
        nr.share[:] = (17 * rank + step) & 0xFF
        nr.mark_done(step)
        if rank == 0:
            nr.wait_all(step, timeout_s=30)
            for r in range(world):
                ok = ok and bool((nr.share_of(r) == ((17 * r + step) & 0xFF)).all())
        dist.barrier()
    if rank == 0:
        try:
            nr.flags[world - 1] = 0
            nr.wait_all(5, timeout_s=0.2)    # a rank that never finishes: a time-out, not a hang
            ok = False
        except TimeoutError:
            pass
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    nr.close()
    dist.destroy_process_group()


def test_node_response_shared_buffer_gloo(tmp_path):
    """world-size-2 check of the e2e response path of bench.py at N > 1: one POSIX shared-memory buffer, every
    rank writes its own share, rank 0 sees all of them once every rank has marked the step"""
    port = _free_port()
    out = tmp_path / "result.txt"
    mp.spawn(_response_worker, args=(2, port, f"pf_test_{port}", str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def _open_response_worker(rank, world, port, out_path):
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = bench.Comm(world, rank, rank, "cpu")
    ok = True
    for round_ in range(2):                  # twice: the segment of the first round is gone before the second
        nr = bench.open_node_response(comm, qbytes=3000 + round_, seg=2000 + 100 * rank)
        ok = ok and nr.shared_data and nr.share.size == 2000 + 100 * rank
        nr.share[:] = rank + 1
        nr.mark_done(0)
        if rank == 0:
            nr.flags[0] = 1
            nr.wait_all(0, timeout_s=30)
            ok = ok and all(bool((nr.share_of(r) == r + 1).all()) for r in range(world))
        dist.barrier()
        nr.close()
        dist.barrier()
    # /dev/shm too small for the whole buffer (rank 0 decides, everybody follows): only the flags are shared, every
    # rank keeps its query copy and its share in private memory; completion still reaches rank 0 through the flags
    bench.shm_has_room = lambda nbytes: False
    nr = bench.open_node_response(comm, qbytes=3000, seg=2000 + 100 * rank)
    ok = ok and not nr.shared_data and nr.share.size == 2000 + 100 * rank and nr.query.size == 3000 and nr.whole.size == nr.FLAGS
    nr.share[:] = rank + 1
    if rank == 0:
        nr.flags[:] = 0
    dist.barrier()
    nr.mark_done(4)
    if rank == 0:
        nr.wait_all(4, timeout_s=30)
        ok = ok and [int(v) for v in nr.flags[:world]] == [5] * world
    dist.barrier()
    nr.close()
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_open_node_response_same_collective_order(tmp_path, world):
    """bench.open_node_response issues the same collectives in the same order on every rank (a broadcast
    between two conditional barriers once deadlocked an 8-GPU run): it must complete under gloo"""
    port = _free_port()
    out = tmp_path / "result.txt"
    mp.spawn(_open_response_worker, args=(world, port, str(out)), nprocs=world, join=True)
    assert out.read_text() == "ok"


def test_stage_guard_emits_the_published_line_when_a_stage_hangs(tmp_path):
    """StageGuard: a hang (or a reported failure) in an optional stage ends the run with exit code 0 and the
    line published before it, `aborted_stage` filled in"""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "g = bench.StageGuard(0, 1)\n"
        "g.publish({'metric': 'm', 'value': 1.5})\n"
        "g.enter('e2e', 0.5)\n"
        "time.sleep(30)\n" % str(root))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.5 and line["aborted_stage"]["stage"] == "e2e"
    code2 = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "g = bench.StageGuard(0, 1)\n"
        "g.publish({'metric': 'm', 'value': 2.5})\n"
        "g.enter('strong', 100)\n"
        "g.abort('strong', 'boom')\n" % str(root))
    r = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 2.5 and "boom" in line["aborted_stage"]["why"]
    # SIGTERM (torchrun tearing the job down): the published line still goes out
    code4 = (
        "import os, signal, sys, time; sys.path.insert(0, %r); import bench\n"
        "g = bench.StageGuard(0, 1)\n"
        "g.publish({'metric': 'm', 'value': 3.5})\n"
        "g.enter('strong', 100)\n"
        "import ctypes, threading\n"
        "threading.Timer(0.5, lambda: os.kill(os.getpid(), signal.SIGTERM)).start()\n"
        "ctypes.CDLL(None).sleep(30)\n" % str(root))   # the main thread sits in C, as it does inside a collective
    r = subprocess.run([sys.executable, "-c", code4], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 3.5 and "SIGTERM" in line["aborted_stage"]["why"]
    # nothing published yet: a non-zero exit code and no line
    code3 = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "g = bench.StageGuard(0, 1, total_limit_s=0.3)\n"
        "time.sleep(30)\n" % str(root))
    r = subprocess.run([sys.executable, "-c", code3], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and not r.stdout.strip()
