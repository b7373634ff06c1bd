"""CPU checks of the boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/prefhetch_b200.h declares.  No compute calls (no GPU here)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from prefhetch_b200 import _capi, build
    build.build()
    return _capi.load()


def test_header_symbols_exported(lib):
    from prefhetch_b200 import _capi
    hdr = (ROOT / "include" / "prefhetch_b200.h").read_text()
    declared = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pf_abi_version() == 2


def test_no_cpu_fallback(lib):
    """without a CUDA device the engine refuses to exist (PF_ERR_CUDA), it never computes on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import prefhetch_b200 as pf
    with pytest.raises(pf.PfError) as ei:
        pf.Engine(128)
    assert ei.value.code == 2 and "no CPU path" in str(ei.value)


def test_product_does_not_touch_oracle():
    """the product tree must not import, link or execute anything under oracle/"""
    for p in (ROOT / "prefhetch_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            txt = p.read_text()
            assert not re.search(r"pf_oracle|import\s+oracle|from\s+oracle|#include\s+\"[^\"]*oracle|libpf_oracle", txt), p


def test_built_for_sm_100a():
    import subprocess
    so = ROOT / "prefhetch_b200" / "libprefhetch_b200.so"
    out = subprocess.run(["cuobjdump", "-lelf", str(so)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_parms_id_is_blake2b_of_the_parameter_words():
    """pf_parms_id (host BLAKE2b, prefhetch_b200/csrc/pf_blake2b.h) against Python's hashlib on SEAL's
    layout {scheme = 1, N, primes..., t}; also exercises multi-block messages (> 128 bytes)."""
    import hashlib
    import struct
    import prefhetch_b200 as pf
    from tests.util import ntt_primes
    cases = [(8192, pf.bfv_default_primes(8192), 16760833), (8192, pf.bfv_default_primes(8192)[:1], 16760833),
             (16384, pf.bfv_default_primes(16384), 16580609), (2048, ntt_primes(2048, 30, 14) + ntt_primes(2048, 31, 6), 65537)]
    for n, primes, t in cases:
        words = [1, n, *primes, t]
        want = struct.unpack("<4Q", hashlib.blake2b(struct.pack(f"<{len(words)}Q", *words), digest_size=32).digest())
        assert pf.parms_id(n, primes, t) == want


def test_seal_stream_inflate_matches_python_zlib(oracle):
    """compr_mode zlib handling of the product (host code, no GPU) against Python's zlib on a stream the
    oracle serialised; corrupt / truncated / zstd streams are refused."""
    import struct
    import zlib
    import prefhetch_b200 as pf
    from tests.util import toy_params
    n, primes, t = toy_params()
    ctx = oracle.Context(n, primes, t)
    rng = np.random.default_rng(1)
    ct = np.stack([[rng.integers(0, q, size=n, dtype=np.uint64) for q in primes[:-1]] for _ in range(2)])
    raw = ctx.ct_save(ct)
    body = zlib.compress(raw[16:], 6)
    z = raw[:5] + b"\x01" + raw[6:8] + struct.pack("<Q", 16 + len(body)) + body
    assert pf.seal_stream_inflate(z) == raw          # inflated, header rewritten to compr none + new size
    assert pf.seal_stream_inflate(raw) == raw        # uncompressed streams pass through
    assert pf.seal_stream_inflate(z + b"tail") == raw  # trailing bytes of a longer buffer are not consumed
    for bad in (z[:-9], z[:5] + b"\x02" + z[6:], z[:5] + b"\x03" + z[6:], z[:40] + bytes(8) + z[48:], b"\x00" * 32):
        with pytest.raises(pf.PfError):      # truncated; a zlib body labelled zstd; an unknown mode; corrupt; not a stream
            pf.seal_stream_inflate(bad)


def test_seal_stream_inflate_zstd(oracle):
    """compr_mode zstd (SEAL's default when built with it): the product binds libzstd.so.1 at run time and decodes
    one-shot frames (content size in the header) and streamed frames (none, what SEAL's ZSTD_compressStream2 loop
    writes); the frames come from pyarrow's own bundled libzstd.  Truncated, corrupt and over-long streams are
    refused; the ceiling holds."""
    import ctypes as C
    import struct
    import prefhetch_b200 as pf
    from prefhetch_b200 import _capi
    from tests.util import have_zstd, toy_params, zstd_stream
    if not have_zstd():
        pytest.skip("pyarrow without the zstd codec: no independent encoder to make vectors with")
    n, primes, t = toy_params()
    ctx = oracle.Context(n, primes, t)
    rng = np.random.default_rng(2)
    ct = np.stack([[rng.integers(0, q, size=n, dtype=np.uint64) for q in primes[:-1]] for _ in range(2)])
    raw = ctx.ct_save(ct)
    for streaming in (False, True):
        z = zstd_stream(raw, streaming)
        assert z[5] == 2 and z[16:20] == bytes([0x28, 0xB5, 0x2F, 0xFD])       # Zstandard frame magic
        assert pf.seal_stream_inflate(z) == raw
        assert pf.seal_stream_inflate(z + b"tail") == raw
        # truncated; a broken frame magic; a block header claiming a reserved block type.  (Flipped payload bytes of a
        # raw block go unnoticed, as in SEAL: it writes frames without the optional content checksum.)
        hdr = 16 + (6 if streaming else 4 + 1 + {0: 1, 1: 2, 2: 4, 3: 8}[z[20] >> 6] + (0 if z[20] & 0x20 else 1))
        for bad in (z[:-5], z[:16] + b"\x00" + z[17:], z[:hdr] + bytes([z[hdr] | 0x06]) + z[hdr + 1:]):
            with pytest.raises(pf.PfError):
                pf.seal_stream_inflate(bad)
    # a zstd bomb: 8 MiB of zeros in a few hundred bytes
    lib = _capi.load()
    payload = bytes(8 << 20)
    z = zstd_stream(bytes(16) + payload, streaming=True)
    z = bytes([0x5E, 0xA1, 0x10, 4, 1, 2, 0, 0]) + z[8:]
    assert len(z) < 4096
    zin = np.frombuffer(z, dtype=np.uint8)
    need, used = C.c_size_t(), C.c_size_t()
    small = np.full(4096 + 64, 0xAB, dtype=np.uint8)
    rc = lib.pf_seal_stream_inflate(zin.ctypes.data_as(C.c_void_p), zin.size, small.ctypes.data_as(C.c_void_p), 4096,
                                    C.byref(need), C.byref(used))
    assert rc == _capi.PF_ERR_CAPACITY and need.value == 16 + len(payload) and (small == 0xAB).all()
    exact = np.zeros(16 + len(payload), dtype=np.uint8)
    rc = lib.pf_seal_stream_inflate(zin.ctypes.data_as(C.c_void_p), zin.size, exact.ctypes.data_as(C.c_void_p), exact.size,
                                    C.byref(need), C.byref(used))
    assert rc == _capi.PF_OK and need.value == exact.size and used.value == len(z) and exact[5] == 0 and not exact[16:].any()


def test_zstd_without_the_library_is_refused():
    """no libzstd on the host (PF_ZSTD_LIB points at nothing): a zstd stream is a format error, not a crash, and
    zlib / uncompressed streams still work.  Separate process: the binding is resolved once per process."""
    import subprocess
    import sys
    from tests.util import have_zstd
    if not have_zstd():
        pytest.skip("pyarrow without the zstd codec")
    code = (
        "import struct, zlib\n"
        "import prefhetch_b200 as pf\n"
        "from tests.util import zstd_stream, zlib_stream\n"
        "raw = bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0]) + struct.pack('<Q', 16 + 4096) + bytes(range(256)) * 16\n"
        "assert pf.seal_stream_inflate(zlib_stream(raw)) == raw\n"
        "try:\n"
        "    pf.seal_stream_inflate(zstd_stream(raw))\n"
        "    print('accepted')\n"
        "except pf.PfError as e:\n"
        "    print('refused', e.code)\n")
    import os
    env = dict(os.environ, PF_ZSTD_LIB="/nonexistent/libzstd.so.1", PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env, cwd=str(ROOT))
    assert r.returncode == 0 and r.stdout.strip().startswith("refused"), r.stdout + r.stderr


def test_seal_stream_inflate_is_bounded():
    """a deflate bomb is refused once it exceeds the caller's capacity: PF_ERR_CAPACITY with the true size
    (below the hard ceiling) and nothing written past the buffer; exact-capacity output still succeeds"""
    import ctypes as C
    import struct
    import zlib
    from prefhetch_b200 import _capi
    lib = _capi.load()
    payload = bytes(3 << 20)
    body = zlib.compress(payload, 9)
    z = bytes([0x5E, 0xA1, 0x10, 4, 1, 1, 0, 0]) + struct.pack("<Q", 16 + len(body)) + body
    zin = np.frombuffer(z, dtype=np.uint8)
    need, used = C.c_size_t(), C.c_size_t()
    small = np.full(4096 + 64, 0xAB, dtype=np.uint8)
    rc = lib.pf_seal_stream_inflate(zin.ctypes.data_as(C.c_void_p), zin.size, small.ctypes.data_as(C.c_void_p), 4096,
                                    C.byref(need), C.byref(used))
    assert rc == _capi.PF_ERR_CAPACITY and need.value == 16 + len(payload)
    assert (small == 0xAB).all()
    exact = np.zeros(16 + len(payload), dtype=np.uint8)
    rc = lib.pf_seal_stream_inflate(zin.ctypes.data_as(C.c_void_p), zin.size, exact.ctypes.data_as(C.c_void_p), exact.size,
                                    C.byref(need), C.byref(used))
    assert rc == _capi.PF_OK and need.value == exact.size and used.value == len(z)
    assert exact[5] == 0 and not exact[16:].any()


def test_handler_json_envelope_cpp():
    """the JSON codec of the handler bodies (prefhetch_b200/host/pf_query_handlers.hpp) on the reference's request
    shapes (ref: src/server/controllers/Query.cc:34-42): compiled and run on the CPU; the handler bodies with an
    engine behind them run on the GPU box (tests/test_gpu_parity.py::test_cpp_handlers_end_to_end)"""
    import subprocess
    exe = ROOT / "prefhetch_b200" / "host" / "pf_handlers_check"
    if not exe.exists():
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_every_entry_point_is_documented():
    """INTEGRATION.md §1 maps every function the header declares to the reference interface it replaces (or says that
    nothing in the reference corresponds); a new entry point has to be added there"""
    import re
    hdr = (ROOT / "include" / "prefhetch_b200.h").read_text()
    funcs = sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr)))
    doc = (ROOT / "INTEGRATION.md").read_text()
    families = [p[:-1] for p in re.findall(r"`(pf_[a-z_]+\*)`", doc)]
    missing = [f for f in funcs if f not in doc and not any(f.startswith(p) for p in families)]
    assert len(funcs) > 50 and not missing, missing
