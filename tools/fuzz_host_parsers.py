#!/usr/bin/env python
"""Mutation fuzz of everything on the host that parses bytes a client or a file supplies, under ASAN / UBSAN.
CPU only.  What it builds and runs (all outputs under /tmp):

  * host/pf_faiss_io.hpp   (pf_faiss_check, g++ -fsanitize=address,undefined): mutated .faiss files; the C++ reader
                           and the Python reader (prefhetch_b200/faiss_io.py) must agree on accept / reject
  * host/pf_json.hpp + pf_client.hpp (pf_client_check): mutated response bodies of POST /coarsesearch-encrypted
  * the SEAL stream parsers of the engine library (pf_seal_stream_inflate, pf_seal_ct_expand: none / zlib / zstd /
    seeded; pf_seal_galois_keys_expand: full / seeded GaloisKeys) through ctypes, the library rebuilt with -fsanitize=address and loaded via PF_LIB

    python tools/fuzz_host_parsers.py            # ~6 min, most of it the ASAN build of the library
Round 2 result: 600 + 300 + 4000 + 1000 mutants, no sanitizer report, no disagreement between the two FAISS readers."""
from __future__ import annotations

import json
import os
import random
import shutil
import struct
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

HOST = ROOT / "prefhetch_b200" / "host"
SAN = ["-std=c++20", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]


def mutate(R: random.Random, good: bytes, header: int) -> bytes:
    b = bytearray(good)
    for _ in range(R.randint(1, 3)):
        op, pos = R.randint(0, 3), R.randrange(len(b))
        if op == 0:
            b[pos] = R.randrange(256)
        elif op == 1:
            del b[pos:pos + R.randint(1, 200)]
        elif op == 2:
            b[pos:pos] = bytes(R.randrange(256) for _ in range(R.randint(1, 16)))
        else:       # a hostile length field somewhere in the header
            pos = R.randrange(0, max(1, min(len(b) - 8, header)))
            b[pos:pos + 8] = R.choice([b"\xff" * 8, (2 ** 63).to_bytes(8, "little"), (2 ** 33).to_bytes(8, "little"), bytes(8)])
    return bytes(b)


def sanitizer_hit(r) -> bool:
    return r.returncode not in (0, 1) or "AddressSanitizer" in r.stderr or "runtime error" in r.stderr


def fuzz_faiss(n=600) -> int:
    from prefhetch_b200 import faiss_io
    exe = "/tmp/pf_faiss_check_asan"
    subprocess.run(["g++", *SAN, "-o", exe, str(HOST / "pf_faiss_check.cpp")], check=True)
    rng = np.random.default_rng(3)
    nlist, d = 16, 128
    ids = [rng.integers(0, 10 ** 6, size=int(k)).astype(np.int64) for k in rng.integers(0, 30, size=nlist)]
    codes = [rng.integers(0, 256, size=(len(x), 32), dtype=np.uint8) for x in ids]
    f = faiss_io.IVFPQFile(d, sum(len(x) for x in ids), nlist, 20, rng.normal(size=(nlist, d)).astype(np.float32), ids, codes,
                           pq_centroids=rng.normal(size=32 * 256 * 4).astype(np.float32))
    faiss_io.write_ivfpq("/tmp/fz.faiss", f)
    good = Path("/tmp/fz.faiss").read_bytes()
    R, bad = random.Random(5), 0
    for it in range(n):
        Path("/tmp/fz_m.faiss").write_bytes(mutate(R, good, 400))
        r = subprocess.run([exe, "/tmp/fz_m.faiss"], capture_output=True, text=True, timeout=60)
        if sanitizer_hit(r):
            print("faiss: sanitizer / crash at mutant", it, r.stderr[-1500:])
            return 1
        try:
            faiss_io.read_ivfpq("/tmp/fz_m.faiss")
            py_ok = True
        except Exception:       # noqa: BLE001
            py_ok = False
        if py_ok != (r.returncode == 0):
            print("faiss: readers disagree at mutant", it, py_ok, r.returncode, r.stderr[-300:])
            bad += 1
    print(f"faiss: {n} mutants, {bad} disagreements")
    return bad


def fuzz_client_json(n=300) -> int:
    import tests.test_client as T
    from oracle import pf_oracle as O
    O.build()
    exe = Path("/tmp/pf_client_check_asan")
    subprocess.run(["g++", *SAN, "-o", str(exe), str(HOST / "pf_client_check.cpp")], check=True)
    tmp = Path("/tmp/fuzz_client_case")
    shutil.rmtree(tmp, ignore_errors=True)
    T.test_client_round_trip_against_the_oracle(O, exe, tmp, 2048, None, 128, 1, 16, 1)     # the whole flow under ASAN
    resp = json.loads((tmp / "response.json").read_text())
    resp["resultsPerQuery"][0] -= 1             # the test leaves an inconsistent envelope behind
    good = json.dumps(resp).encode()
    R = random.Random(1)
    for it in range(n):
        (tmp / "response.json").write_bytes(mutate(R, good, 64))
        r = subprocess.run([str(exe), "respond", str(tmp)], capture_output=True, text=True)
        if sanitizer_hit(r):
            print("client json: sanitizer / crash at mutant", it, r.stderr[-1500:])
            return 1
    print(f"client json: {n} mutants clean")
    return 0


def fuzz_seal_streams(n=4000) -> int:
    if os.environ.get("PF_LIB") != "/tmp/libpf_asan.so":       # re-exec with the ASAN build of the library preloaded
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-g", "-std=c++17", "-shared", "-Xcompiler",
                        "-fPIC,-fsanitize=address,-fno-omit-frame-pointer", "-o", "/tmp/libpf_asan.so",
                        str(ROOT / "prefhetch_b200" / "csrc" / "pf_engine.cu"), "-lz", "-ldl"], check=True)
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", PF_LIB="/tmp/libpf_asan.so")
        return subprocess.run([sys.executable, __file__, "--seal-child"], env=env).returncode
    import prefhetch_b200 as pf
    from oracle import pf_oracle as O
    from tests.util import toy_params, zlib_stream, zstd_stream
    O.build()
    nn, primes, t = toy_params()
    ctx = O.Context(nn, primes, t)
    rng = np.random.default_rng(1)
    sk, seed = ctx.keygen(3), rng.bytes(64)
    ct = ctx.encrypt_seeded(sk, ctx.encode(rng.integers(0, t, size=nn, dtype=np.uint64)), 7, seed)
    full, seeded = ctx.ct_save(ct), ctx.ct_save_seeded(ct, seed)
    goods = [full, seeded, zlib_stream(full), zlib_stream(seeded), zstd_stream(full), zstd_stream(seeded, True)]
    cap = 113 + 2 * (len(primes) - 1) * nn * 8
    R, acc = random.Random(9), 0
    for _ in range(n):
        b = mutate(R, R.choice(goods), 130)
        for fn in (pf.seal_stream_inflate, lambda x: pf.seal_ct_expand(x, nn, primes[:-1])):
            try:
                out = fn(b)
                acc += 1
                assert fn is pf.seal_stream_inflate or len(out) == cap, "pf_seal_ct_expand returned something that is not a full ciphertext stream"
            except pf.PfError:
                pass
    print(f"seal streams: {n} mutants x 2 parsers clean under ASAN ({acc} accepted)")
    # GaloisKeys streams (full and seeded, as the C++ client writes them) through pf_seal_galois_keys_expand
    import tests.test_client as T
    exe = HOST / "pf_client_check"
    tmp = Path("/tmp/fuzz_gk_case")
    shutil.rmtree(tmp, ignore_errors=True)
    kn, kprimes, kt = 1024, T.ntt_primes(1024, 40, 2) + T.ntt_primes(1024, 41, 1), T.ntt_primes(1024, 20, 1)[0]
    T.write_case(tmp, kn, kprimes, kt, 64, 1, 16, np.zeros((1, 64), dtype=np.int64), 1, 1, bytes(range(64)))
    subprocess.run([str(exe), "keygen", str(tmp)], check=True, capture_output=True)
    gks = [(tmp / "galois_keys.bin").read_bytes(), (tmp / "galois_keys_full.bin").read_bytes()]
    gks += [zlib_stream(gks[0]), zstd_stream(gks[0], True)]
    want, acc = gks[1], 0
    for _ in range(n // 4):
        b = mutate(R, R.choice(gks), 80 + 8 * kn)
        try:
            out = pf.seal_galois_keys_expand(b, kn, kprimes)
            acc += 1
            assert len(out) <= 2 * len(want) + 4096
        except pf.PfError:
            pass
    print(f"galois keys: {n // 4} mutants clean under ASAN ({acc} accepted)")
    return 0


if __name__ == "__main__":
    if "--seal-child" in sys.argv:
        sys.exit(fuzz_seal_streams())
    struct.calcsize("<Q")
    sys.exit(fuzz_faiss() or fuzz_client_json() or fuzz_seal_streams())
