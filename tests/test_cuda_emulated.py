"""The `-m gpu` parity suite dry-run on the CPU: the product's own CUDA sources (prefhetch_b200/csrc, kernels and
host code alike) rewritten to plain C++ and executed by tests/cuda_emul (a small interpreter of the CUDA execution
model: blocks on host threads, threads as fibers, counted barriers, synchronous streams) behind the same C ABI, with
tests/test_gpu_parity.py run against that library in a child process (PF_LIB / LD_LIBRARY_PATH).

TEST INFRASTRUCTURE: nothing in the package knows about the emulated library, it is built into a scratch directory
and it is no CPU path of the product (a B200 is ~10^4 x faster; `import prefhetch_b200` on a machine without the
sm_100a build still fails).  What a green run shows: the kernels' integer / FP64 arithmetic, indexing and barrier
structure, every host-side line of pf_engine.cu, the ctypes layer and the tests' own expectations agree with the
oracle — before any of it reaches a GPU.  What it cannot show: timing, inter-warp memory-model races, asynchronous
stream ordering (streams are synchronous here), PTX / SASS code generation."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

# the longest cases (33 s each: 64 queries x 1100 results at N = 8192; 13 s: N = 16384 variants) are kept to one
# representative so that the CPU suite stays within minutes; PF_EMUL_FULL=1 runs everything (about 2.5 minutes)
DESELECT = ["test_encrypted_search_bench_shape[0]", "test_kernel_variants_bit_identical[16384-16-2]"]


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    sys.path.insert(0, str(ROOT / "tests" / "cuda_emul"))
    import build_emul
    from prefhetch_b200 import build as b
    b.build()                       # the host check programs link against the product library's name
    if not (ROOT / "prefhetch_b200" / "host" / "pf_roundtrip_example").exists():
        import __graft_entry__ as ge
        ge.build()
    out = tmp_path_factory.mktemp("cuda_emul")
    so = build_emul.build(out)
    ld = out / "ld"
    ld.mkdir()
    (ld / "libprefhetch_b200.so").symlink_to(so)      # what the C++ check programs resolve (LD_LIBRARY_PATH before RUNPATH)
    env = dict(os.environ, PF_LIB=str(so), LD_LIBRARY_PATH=str(ld) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return so, env


def test_rewrite_knows_every_construct():
    """the textual rewrite covers every launch, dynamic shared-memory declaration and inline-PTX statement of the
    product sources (it aborts on one it does not know), and leaves no CUDA-only syntax behind"""
    sys.path.insert(0, str(ROOT / "tests" / "cuda_emul"))
    import build_emul
    nlaunch = 0
    for p in sorted((ROOT / "prefhetch_b200" / "csrc").iterdir()):
        if p.suffix in (".cu", ".cuh", ".h"):
            src = p.read_text()
            out = build_emul.rewrite(src, p.name)
            code = re.sub(r"//[^\n]*", "", out)
            assert not re.search(r"\basm\s*(volatile\s*)?\(", code) and "<<<" not in code, p.name
            nlaunch += out.count("pf_emul::launch(")
    assert nlaunch >= 40


def test_gpu_suite_on_the_emulated_device(emul):
    so, env = emul
    cmd = [sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_gpu_parity.py"), "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"]
    if os.environ.get("PF_EMUL_FULL") != "1":
        for d in DESELECT:
            cmd += ["--deselect", f"tests/test_gpu_parity.py::{d}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd=str(ROOT))
    tail = (r.stdout + r.stderr)[-6000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout, tail


def test_smoke_on_the_emulated_device(emul):
    """__graft_entry__.smoke() — the driver's first call on the GPU box — end to end on the emulated device"""
    so, env = emul
    # threads of a block resumed in a random order every scheduling pass: results must not depend on it
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], capture_output=True, text=True,
                       timeout=900, env=dict(env, PF_EMUL_ORDER="random:3"), cwd=str(ROOT))
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]


def test_bench_on_the_emulated_device(emul):
    """bench.py's whole N = 1 flow — setup, timed loops, recall, the pipelined e2e submit / collect, the parity
    self-check on real encryptions, the CPU baseline, the JSON line — through the REAL Engine over the emulated
    device (PF_BENCH_DRYRUN=emul: no-op torch streams; tests/test_bench_dryrun.py covers the multi-rank control flow
    with a stub engine).  The numbers mean nothing; the line's structure and the parity count do."""
    import json
    so, env = emul
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", "siftsmall_nlist100_nprobe8", "--nq", "4", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900, env=dict(env, PF_BENCH_DRYRUN="emul"), cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    line = json.loads(lines[0])
    assert line.get("aborted_stage") is None
    assert line["parity_checked"] == 32 and line["parity"]["results_in_batch"] == 32
    assert line["gpu_launches"] > 0 and line["value"] > 0 and line["roofline"]["achieved"] > 0
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 4 * (113 + 2 * 4 * 8192 * 8) and e2e["d2h_bytes_per_step"] >= 32 * 2 * 8192 * 8
    es = line["e2e_seeded_requests"]     # the request of a symmetric-key SEAL client: half the upload, c1 drawn on the device
    assert es["value"] > 0 and es["d2h_bytes_per_step"] == e2e["d2h_bytes_per_step"]
    assert es["h2d_bytes_per_step"] < 0.51 * e2e["h2d_bytes_per_step"] + 4096 and "seeded" in es["request"]
    assert line["cpu_baseline"]["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["recall_at_10"] is not None and line["config"]["queries_per_step"] == 4
    assert set(line["phases_ms_per_step"]) == {"coarse", "to_ntt", "rotate", "mac", "intt"}


@pytest.mark.skipif(os.environ.get("PF_EMUL_ASAN") != "1", reason="8.5 minutes: PF_EMUL_ASAN=1 runs the memcheck (result recorded in profiles/README.md)")
def test_gpu_suite_memcheck_under_asan(tmp_path):
    """the whole -m gpu suite with the emulated device built under AddressSanitizer: device allocations and the
    dynamic shared memory of a launch are exact-size heap blocks, so any kernel (or host) access outside them
    aborts the run with a report — compute-sanitizer memcheck's job, on the CPU"""
    sys.path.insert(0, str(ROOT / "tests" / "cuda_emul"))
    import build_emul
    so = build_emul.build(tmp_path, asan=True)
    ld = tmp_path / "ld"
    ld.mkdir()
    (ld / "libprefhetch_b200.so").symlink_to(so)
    env = dict(os.environ, PF_LIB=str(so), LD_PRELOAD=build_emul.asan_runtime(), LD_LIBRARY_PATH=str(ld),
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_gpu_parity.py"), "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"],
                       capture_output=True, text=True, timeout=3000, env=env, cwd=str(ROOT))
    assert r.returncode == 0 and "AddressSanitizer" not in (r.stdout + r.stderr), (r.stdout + r.stderr)[-6000:]


def _torchrun_bench(env, n, port, extra_env=None):
    import json
    e = dict(env, PF_BENCH_DRYRUN="emul", PF_EMUL_IPC="1", PF_EMUL_DEVICES="8", PF_EMUL_THREADS=str(max(1, 8 // n)),
             PF_BENCH_DRYRUN_NB="6000", PF_BENCH_DRYRUN_NQ="8", PF_BENCH_DRYRUN_NPROBE="2", PF_BENCH_DRYRUN_NLIST="24")
    e.update(extra_env or {})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(ROOT / "bench.py"), "--gpus", str(n), "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=1500, env=e, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("n", [2] + ([4, 8] if os.environ.get("PF_EMUL_FULL") == "1" else []))
def test_multi_rank_bench_on_emulated_devices(emul, n):
    """the driver's `bench.py --gpus N` default run with the REAL engine on every rank: emulated devices whose
    allocations live in POSIX shared memory (PF_EMUL_IPC=1), so the peer-buffer gather (IPC handles, copy into rank
    0's buffer, arrival / ack flag kernels) moves real result ciphertexts between processes and `gather_verified`
    compares real checksums; then the e2e pass through the one shared response buffer, the strong-scaling record on
    its L x Q rank grid with its 1-GPU reference, and at 8 ranks the configs[4] stage.  gloo instead of NCCL,
    synchronous streams: the protocol and the bookkeeping are what is checked."""
    so, env = emul
    line = _torchrun_bench(env, n, 29700 + n)
    assert line.get("aborted_stage") is None and line["n_gpus"] == n and line["scaling"] == "weak"
    gv = line["gather_verified"]
    assert gv["ranks"] == n - 1 and gv["results"] > 0 and "mismatch_ranks" not in gv
    assert line["e2e"]["value"] > 0 and "one host buffer per node" in line["e2e"]["response"]
    st = line["strong"]
    assert st["scaling"] == "strong" and st["value"] > 0 and st["value_1gpu"] > 0 and st["e2e"]["value"] > 0
    assert st["gather_verified"]["ranks"] == n - 1 and "mismatch_ranks" not in st["gather_verified"]
    assert st["config"]["parallelism"].startswith({2: "grid 2 list shards x 1", 4: "grid 2 list shards x 2", 8: "grid 2 list shards x 4"}[n])
    if n == 8:
        c4 = line["configs4"]
        assert c4["value"] > 0 and c4["gather_verified"]["ranks"] == 7 and c4["e2e"]["value"] > 0
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("pf_emul_") or f.startswith("pf_bench_297")]
