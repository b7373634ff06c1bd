"""ctypes binding of the CPU oracle (oracle/pf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``prefhetch_b200/`` imports it.
PARITY UNPINNED by the reference (see pf_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libpf_oracle.so"

# SEAL: util/globals.cpp default_coeff_modulus_128 (CoeffModulus::BFVDefault); SURVEY.md App. A.1
BFV_DEFAULT_PRIMES = {
    4096: [0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001],
    8192: [0x7FFFFFD8001, 0x7FFFFFC8001, 0xFFFFFFFC001, 0xFFFFFF6C001, 0xFFFFFEBC001],
    16384: [
        0xFFFFFFFD8001, 0xFFFFFFFA0001, 0xFFFFFFF00001, 0x1FFFFFFF68001, 0x1FFFFFFF50001,
        0x1FFFFFFEE8001, 0x1FFFFFFEA0001, 0x1FFFFFFE88001, 0x1FFFFFFE48001,
    ],
}
# SEAL: PlainModulus::Batching(N, bits) — largest `bits`-bit prime = 1 mod 2N; SURVEY.md App. A.2
BATCHING_T = {(8192, 24): 16760833, (8192, 27): 133857281, (16384, 24): 16580609, (16384, 27): 133857281,
              (4096, 24): 16760833, (4096, 20): 1032193}


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc via oracle/Makefile (outputs only into oracle/_build)."""
    srcs = [_HERE / "pf_oracle.c", _HERE / "pf_oracle_pipeline.c", _HERE / "pf_oracle_seeded.c", _HERE / "pf_oracle.h"]
    if force or not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(_LIB_PATH))
        _declare(_lib)
    return _lib


u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class Modulus(C.Structure):
    _fields_ = [("q", C.c_uint64), ("ratio", C.c_uint64 * 2)]


class Layout(C.Structure):
    _fields_ = [("n", C.c_uint64), ("d", C.c_uint32), ("d_pad", C.c_uint32), ("m", C.c_uint32), ("g", C.c_uint32),
                ("dc", C.c_uint32), ("R", C.c_uint32), ("K", C.c_uint32), ("C", C.c_uint32)]


def _declare(l):
    vp = C.c_void_p
    l.pfo_modulus_init.argtypes = [C.POINTER(Modulus), C.c_uint64]
    l.pfo_modulus_init.restype = C.c_int
    for name in ("pfo_barrett64",):
        getattr(l, name).argtypes = [C.c_uint64, C.POINTER(Modulus)]
        getattr(l, name).restype = C.c_uint64
    l.pfo_barrett128.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(Modulus)]
    l.pfo_barrett128.restype = C.c_uint64
    for name in ("pfo_mulmod", "pfo_powmod"):
        getattr(l, name).argtypes = [C.c_uint64, C.c_uint64, C.POINTER(Modulus)]
        getattr(l, name).restype = C.c_uint64
    l.pfo_invmod.argtypes = [C.c_uint64, C.POINTER(Modulus)]
    l.pfo_invmod.restype = C.c_uint64
    l.pfo_shoup.argtypes = [C.c_uint64, C.c_uint64]
    l.pfo_shoup.restype = C.c_uint64
    l.pfo_minimal_primitive_root.argtypes = [C.c_uint64, C.POINTER(Modulus)]
    l.pfo_minimal_primitive_root.restype = C.c_uint64
    l.pfo_context_create.argtypes = [C.c_uint64, u64p, C.c_int, C.c_uint64]
    l.pfo_context_create.restype = vp
    l.pfo_context_destroy.argtypes = [vp]
    l.pfo_tables.argtypes = [vp, C.c_int]
    l.pfo_tables.restype = vp
    l.pfo_ntt_fwd.argtypes = [u64p, vp]
    l.pfo_ntt_inv.argtypes = [u64p, vp]
    l.pfo_batch_encode.argtypes = [vp, u64p, C.c_uint64, u64p]
    l.pfo_batch_decode.argtypes = [vp, u64p, u64p]
    l.pfo_plain_to_ntt.argtypes = [vp, u64p, u64p]
    l.pfo_add_plain_scaled.argtypes = [vp, u64p, u64p]
    l.pfo_keygen.argtypes = [vp, C.c_uint64, u64p]
    l.pfo_galois_keygen.argtypes = [vp, u64p, C.c_uint32, C.c_uint64, u64p]
    l.pfo_encrypt_symmetric.argtypes = [vp, u64p, u64p, C.c_uint64, u64p]
    l.pfo_decrypt.argtypes = [vp, u64p, u64p, u64p]
    l.pfo_decrypt.restype = C.c_int
    l.pfo_ct_to_ntt.argtypes = [vp, u64p, C.c_int]
    l.pfo_ct_from_ntt.argtypes = [vp, u64p, C.c_int]
    l.pfo_multiply_plain_ntt.argtypes = [vp, u64p, u64p, u64p]
    l.pfo_add.argtypes = [vp, u64p, u64p]
    l.pfo_mac_plain_ntt.argtypes = [vp, u64p, u64p, C.c_size_t, C.c_int, u64p]
    l.pfo_galois_elt_from_step.argtypes = [vp, C.c_int]
    l.pfo_galois_elt_from_step.restype = C.c_uint32
    l.pfo_apply_galois.argtypes = [vp, u64p, C.c_uint32, C.c_int, u64p]
    l.pfo_apply_galois_ntt.argtypes = [vp, u64p, C.c_uint32, u64p]
    l.pfo_galois_ntt_table.argtypes = [vp, C.c_uint32, u32p]
    l.pfo_switch_key.argtypes = [vp, u64p, u64p, u64p]
    l.pfo_apply_galois_ct.argtypes = [vp, u64p, C.c_uint32, u64p]
    l.pfo_mod_switch_next.argtypes = [vp, u64p, C.c_int, u64p]
    l.pfo_ct_save_size.argtypes = [C.c_uint64, C.c_int, C.c_int]
    l.pfo_ct_save_size.restype = C.c_size_t
    l.pfo_ct_save.argtypes = [u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, u64p, C.POINTER(C.c_uint8)]
    l.pfo_ct_save.restype = C.c_size_t
    l.pfo_ct_load.argtypes = [C.POINTER(C.c_uint8), C.c_size_t, u64p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                              C.POINTER(C.c_int), u64p, u64p, C.c_size_t]
    l.pfo_ct_load.restype = C.c_size_t
    l.pfo_vecs_read.argtypes = [C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(vp)]
    l.pfo_vecs_read.restype = C.c_int
    l.pfo_coarse_quantize.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, f32p, f32p, C.c_size_t, i64p, f32p]
    l.pfo_l2sqr_ref.argtypes = [f32p, f32p, C.c_size_t]
    l.pfo_l2sqr_ref.restype = C.c_float
    l.pfo_search_lists_plain.argtypes = [C.c_size_t, C.c_size_t, f32p, i64p, C.c_size_t, i64p, i64p, f32p, f32p,
                                         i64p, C.c_size_t, C.POINTER(C.c_size_t)]
    l.pfo_search_lists_plain.restype = C.c_size_t
    u8p_ = C.POINTER(C.c_uint8)
    l.pfo_search_lists_pq.argtypes = [C.c_size_t, C.c_size_t, f32p, i64p, C.c_size_t, f32p, i64p, i64p, C.c_size_t, f32p, u8p_,
                                      f32p, i64p, C.c_size_t, C.POINTER(C.c_size_t)]
    l.pfo_search_lists_pq.restype = C.c_size_t
    l.pfo_pq_encode_residuals.argtypes = [C.c_size_t, C.c_size_t, f32p, i64p, f32p, C.c_size_t, f32p, u8p_]
    l.pfo_pq_encode_residuals.restype = None
    dp = C.POINTER(C.c_double)
    l.pfo_recall.argtypes = [C.c_size_t, C.c_size_t, i64p, C.c_size_t, i32p, dp, dp, dp, dp, dp]
    l.pfo_layout_init.argtypes = [C.POINTER(Layout), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    l.pfo_layout_init.restype = C.c_int
    l.pfo_layout_query_slots.argtypes = [C.POINTER(Layout), C.c_uint64, i64p, C.c_uint32, u64p]
    l.pfo_layout_slot.argtypes = [C.POINTER(Layout), C.c_uint32, C.c_uint32]
    l.pfo_layout_slot.restype = C.c_uint32
    l.pfo_layout_diag_slots.argtypes = [C.POINTER(Layout), C.c_uint64, i32p, C.c_uint32, C.c_uint32, C.c_uint32, u64p]
    l.pfo_layout_norm_slots.argtypes = [C.POINTER(Layout), C.c_uint64, i32p, C.c_uint32, u64p]
    l.pfo_encode_block.argtypes = [vp, C.POINTER(Layout), i32p, C.c_uint32, u64p, u64p]
    l.pfo_rotate_query_set.argtypes = [vp, C.POINTER(Layout), u64p, C.POINTER(u64p), C.c_int, u64p]
    l.pfo_block_distance.argtypes = [vp, C.POINTER(Layout), u64p, u64p, u64p, u64p]
    l.pfo_max_threads.restype = C.c_int
    l.pfo_encode_blocks.argtypes = [vp, C.POINTER(Layout), C.c_size_t, i32p, i64p, u32p, u64p, u64p, C.c_int]
    l.pfo_search_pairs.argtypes = [vp, C.POINTER(Layout), C.c_size_t, u64p, C.POINTER(u64p), C.c_int, C.c_size_t,
                                   i32p, i64p, u64p, u64p, u64p, u64p, C.c_int, dp]
    u8p = C.POINTER(C.c_uint8)
    l.pfo_blake2b_param.argtypes = [u8p, u8p, C.c_size_t, u8p, C.c_size_t, u8p, C.c_size_t]
    l.pfo_blake2xb.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, u8p, C.c_size_t]
    l.pfo_seal_sample_poly_uniform.argtypes = [vp, C.c_int, u8p, u64p]
    l.pfo_encrypt_symmetric_seeded.argtypes = [vp, u64p, u64p, C.c_uint64, u8p, u64p]
    l.pfo_ct_save_seeded_size.argtypes = [C.c_uint64, C.c_int]
    l.pfo_ct_save_seeded_size.restype = C.c_size_t
    l.pfo_ct_save_seeded.argtypes = [u64p, C.c_uint64, C.c_int, u64p, u8p, C.c_uint8, u8p]
    l.pfo_ct_save_seeded.restype = C.c_size_t
    l.pfo_search_pairs_ms.argtypes = [vp, C.POINTER(Layout), C.c_size_t, u64p, C.POINTER(u64p), C.c_int, C.c_size_t,
                                      i32p, i64p, u64p, u64p, u64p, C.c_int, u64p, C.c_int, dp]


def _p(a: np.ndarray, typ):
    assert a.flags["C_CONTIGUOUS"], "array must be contiguous"
    return a.ctypes.data_as(typ)


def modulus(q: int) -> Modulus:
    m = Modulus()
    if lib().pfo_modulus_init(C.byref(m), q):
        raise ValueError(f"bad modulus {q}")
    return m


class Context:
    """BFV context: N, coefficient primes (last = special prime), plain modulus t."""

    def __init__(self, n: int, primes, t: int):
        self.n, self.primes, self.t = int(n), [int(p) for p in primes], int(t)
        self.k, self.L = len(self.primes), len(self.primes) - 1
        arr = (C.c_uint64 * self.k)(*self.primes)
        self.h = lib().pfo_context_create(self.n, arr, self.k, self.t)
        if not self.h:
            raise ValueError("pfo_context_create failed (primes must be = 1 mod 2N)")
        self.ctw = 2 * self.L * self.n

    def __del__(self):
        if getattr(self, "h", None):
            lib().pfo_context_destroy(self.h)
            self.h = None

    # --- NTT ---
    def ntt_fwd(self, a: np.ndarray, limb: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().pfo_ntt_fwd(_p(a, u64p), lib().pfo_tables(self.h, limb))
        return a

    def ntt_inv(self, a: np.ndarray, limb: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().pfo_ntt_inv(_p(a, u64p), lib().pfo_tables(self.h, limb))
        return a

    # --- encoder ---
    def encode(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.zeros(self.n, dtype=np.uint64)
        lib().pfo_batch_encode(self.h, _p(v, u64p), len(v), _p(out, u64p))
        return out

    def decode(self, plain: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        out = np.zeros(self.n, dtype=np.uint64)
        lib().pfo_batch_decode(self.h, _p(p, u64p), _p(out, u64p))
        return out

    def plain_to_ntt(self, plain: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        out = np.zeros((self.L, self.n), dtype=np.uint64)
        lib().pfo_plain_to_ntt(self.h, _p(p, u64p), _p(out, u64p))
        return out

    def add_plain_scaled(self, plain: np.ndarray, poly0: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        out = np.ascontiguousarray(poly0, dtype=np.uint64).copy()
        lib().pfo_add_plain_scaled(self.h, _p(p, u64p), _p(out, u64p))
        return out

    # --- keys / enc / dec ---
    def keygen(self, seed: int) -> np.ndarray:
        sk = np.zeros((self.k, self.n), dtype=np.uint64)
        lib().pfo_keygen(self.h, seed, _p(sk, u64p))
        return sk

    def galois_keygen(self, sk: np.ndarray, elt: int, seed: int) -> np.ndarray:
        key = np.zeros((self.L, 2, self.k, self.n), dtype=np.uint64)
        lib().pfo_galois_keygen(self.h, _p(sk, u64p), elt, seed, _p(key, u64p))
        return key

    def encrypt(self, sk: np.ndarray, plain: np.ndarray, seed: int) -> np.ndarray:
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        ct = np.zeros((2, self.L, self.n), dtype=np.uint64)
        lib().pfo_encrypt_symmetric(self.h, _p(sk, u64p), _p(p, u64p), seed, _p(ct, u64p))
        return ct

    def decrypt(self, sk: np.ndarray, ct: np.ndarray):
        c = np.ascontiguousarray(ct, dtype=np.uint64)
        plain = np.zeros(self.n, dtype=np.uint64)
        budget = lib().pfo_decrypt(self.h, _p(sk, u64p), _p(c, u64p), _p(plain, u64p))
        return plain, budget

    # --- SEAL seeded ciphertexts (pf_oracle_seeded.c) ---
    def sample_poly_uniform(self, seed: bytes) -> np.ndarray:
        """util::sample_poly_uniform over the L data primes from a Blake2xbPRNG with this 64-byte seed"""
        out = np.zeros((self.L, self.n), dtype=np.uint64)
        sb = np.frombuffer(seed, dtype=np.uint8).copy()
        lib().pfo_seal_sample_poly_uniform(self.h, self.L, _p(sb, C.POINTER(C.c_uint8)), _p(out, u64p))
        return out

    def encrypt_seeded(self, sk: np.ndarray, plain: np.ndarray, noise_seed: int, seed: bytes) -> np.ndarray:
        """symmetric encryption whose c1 is the expansion of `seed` (what a seeded save / load round-trips)"""
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        ct = np.zeros((2, self.L, self.n), dtype=np.uint64)
        sb = np.frombuffer(seed, dtype=np.uint8).copy()
        lib().pfo_encrypt_symmetric_seeded(self.h, _p(sk, u64p), _p(p, u64p), noise_seed, _p(sb, C.POINTER(C.c_uint8)), _p(ct, u64p))
        return ct

    def ct_save_seeded(self, ct: np.ndarray, seed: bytes, parms_id=(0, 0, 0, 0), prng_type: int = 1) -> bytes:
        """Serializable<Ciphertext>::save of a seeded ciphertext: c0 and the PRNG seed instead of c1"""
        buf = np.zeros(lib().pfo_ct_save_seeded_size(self.n, self.L), dtype=np.uint8)
        pid = (C.c_uint64 * 4)(*parms_id)
        sb = np.frombuffer(seed, dtype=np.uint8).copy()
        c0 = np.ascontiguousarray(ct[0], dtype=np.uint64)
        w = lib().pfo_ct_save_seeded(_p(c0, u64p), self.n, self.L, pid, _p(sb, C.POINTER(C.c_uint8)), prng_type,
                                     _p(buf, C.POINTER(C.c_uint8)))
        return buf[:w].tobytes()

    # --- evaluator ---
    def ct_to_ntt(self, ct: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        lib().pfo_ct_to_ntt(self.h, _p(c, u64p), 2)
        return c

    def ct_from_ntt(self, ct: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        lib().pfo_ct_from_ntt(self.h, _p(c, u64p), 2)
        return c

    def multiply_plain_ntt(self, ct: np.ndarray, pt: np.ndarray) -> np.ndarray:
        out = np.zeros((2, self.L, self.n), dtype=np.uint64)
        lib().pfo_multiply_plain_ntt(self.h, _p(np.ascontiguousarray(ct), u64p), _p(np.ascontiguousarray(pt), u64p),
                                     _p(out, u64p))
        return out

    def add(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        out = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().pfo_add(self.h, _p(out, u64p), _p(np.ascontiguousarray(b), u64p))
        return out

    def mac_plain_ntt(self, cts: np.ndarray, pts: np.ndarray) -> np.ndarray:
        """cts [K][2][L][n], pts [K][L][n] -> [2][L][n]"""
        K = cts.shape[0]
        out = np.zeros((2, self.L, self.n), dtype=np.uint64)
        lib().pfo_mac_plain_ntt(self.h, _p(np.ascontiguousarray(cts), u64p), _p(np.ascontiguousarray(pts), u64p),
                                self.L * self.n, K, _p(out, u64p))
        return out

    def galois_elt(self, step: int) -> int:
        return lib().pfo_galois_elt_from_step(self.h, step)

    def apply_galois(self, poly: np.ndarray, elt: int, limb: int) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint64)
        lib().pfo_apply_galois(self.h, _p(np.ascontiguousarray(poly, dtype=np.uint64), u64p), elt, limb, _p(out, u64p))
        return out

    def apply_galois_ntt(self, poly: np.ndarray, elt: int) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint64)
        lib().pfo_apply_galois_ntt(self.h, _p(np.ascontiguousarray(poly, dtype=np.uint64), u64p), elt, _p(out, u64p))
        return out

    def galois_ntt_table(self, elt: int) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint32)
        lib().pfo_galois_ntt_table(self.h, elt, _p(out, u32p))
        return out

    def switch_key(self, ct: np.ndarray, target: np.ndarray, key: np.ndarray) -> np.ndarray:
        out = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        lib().pfo_switch_key(self.h, _p(out, u64p), _p(np.ascontiguousarray(target), u64p),
                             _p(np.ascontiguousarray(key), u64p))
        return out

    def rotate_rows(self, ct: np.ndarray, step: int, key: np.ndarray) -> np.ndarray:
        """Evaluator::rotate_rows on a coefficient-form ciphertext with the key for 3^step."""
        out = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        lib().pfo_apply_galois_ct(self.h, _p(out, u64p), self.galois_elt(step), _p(np.ascontiguousarray(key), u64p))
        return out

    def mod_switch_next(self, ct: np.ndarray) -> np.ndarray:
        Lin = ct.shape[1]
        out = np.zeros((2, Lin - 1, self.n), dtype=np.uint64)
        lib().pfo_mod_switch_next(self.h, _p(np.ascontiguousarray(ct), u64p), Lin, _p(out, u64p))
        return out

    # --- wire format ---
    def ct_save(self, ct: np.ndarray, is_ntt: bool = False, parms_id=(0, 0, 0, 0)) -> bytes:
        size, L, n = ct.shape
        buf = np.zeros(lib().pfo_ct_save_size(n, L, size), dtype=np.uint8)
        pid = (C.c_uint64 * 4)(*parms_id)
        w = lib().pfo_ct_save(_p(np.ascontiguousarray(ct, dtype=np.uint64), u64p), n, L, size, int(is_ntt), pid,
                              _p(buf, C.POINTER(C.c_uint8)))
        return buf[:w].tobytes()

    @staticmethod
    def ct_load(data: bytes):
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        n, L, size, is_ntt = C.c_uint64(), C.c_int(), C.c_int(), C.c_int()
        pid = (C.c_uint64 * 4)()
        out = np.zeros(len(data) // 8 + 1, dtype=np.uint64)
        used = lib().pfo_ct_load(_p(buf, C.POINTER(C.c_uint8)), len(buf), C.byref(n), C.byref(L), C.byref(size),
                                 C.byref(is_ntt), pid, _p(out, u64p), out.size)
        if not used:
            raise ValueError("pfo_ct_load failed")
        words = size.value * L.value * n.value
        return out[:words].reshape(size.value, L.value, n.value), bool(is_ntt.value), tuple(pid), used


class LayoutPlan:
    """Generalised-diagonal layout (pfo_layout)."""

    def __init__(self, n: int, d: int, m: int = 1, g: int = 8):
        self.s = Layout()
        if lib().pfo_layout_init(C.byref(self.s), n, d, m, g):
            raise ValueError("bad layout parameters")
        for f, _ in Layout._fields_:
            setattr(self, f, getattr(self.s, f))

    def query_slots(self, t: int, q, a: int = 0) -> np.ndarray:
        qq = np.ascontiguousarray(q, dtype=np.int64)
        out = np.zeros(self.n, dtype=np.uint64)
        lib().pfo_layout_query_slots(C.byref(self.s), t, _p(qq, i64p), a, _p(out, u64p))
        return out

    def slot(self, u: int, j: int) -> int:
        return lib().pfo_layout_slot(C.byref(self.s), u, j)

    def slot_table(self) -> np.ndarray:
        """[C][g] slot indices"""
        return np.array([[self.slot(u, j) for j in range(self.g)] for u in range(self.C)], dtype=np.int64)

    def diag_slots(self, t: int, xs: np.ndarray, a: int, r: int) -> np.ndarray:
        x = np.ascontiguousarray(xs, dtype=np.int32)
        out = np.zeros(self.n, dtype=np.uint64)
        lib().pfo_layout_diag_slots(C.byref(self.s), t, _p(x, i32p), x.shape[0], a, r, _p(out, u64p))
        return out

    def norm_slots(self, t: int, xs: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(xs, dtype=np.int32)
        out = np.zeros(self.n, dtype=np.uint64)
        lib().pfo_layout_norm_slots(C.byref(self.s), t, _p(x, i32p), x.shape[0], _p(out, u64p))
        return out


def encode_block(ctx: Context, lay: LayoutPlan, xs: np.ndarray):
    x = np.ascontiguousarray(xs, dtype=np.int32).reshape(-1, lay.d)
    diag = np.zeros((lay.K, ctx.L, ctx.n), dtype=np.uint64)
    norm = np.zeros((ctx.L, ctx.n), dtype=np.uint64)
    lib().pfo_encode_block(ctx.h, C.byref(lay.s), _p(x, i32p), x.shape[0], _p(diag, u64p), _p(norm, u64p))
    return diag, norm


def encode_blocks(ctx: Context, lay: LayoutPlan, xs: np.ndarray, block_vec_offset, nvec, nthreads: int = 0):
    """xs [ntotal][d] int32 in list order; blocks described by (offset, count)."""
    x = np.ascontiguousarray(xs, dtype=np.int32)
    off = np.ascontiguousarray(block_vec_offset, dtype=np.int64)
    nv = np.ascontiguousarray(nvec, dtype=np.uint32)
    nb = len(off)
    diag = np.zeros((nb, lay.K, ctx.L, ctx.n), dtype=np.uint64)
    norm = np.zeros((nb, ctx.L, ctx.n), dtype=np.uint64)
    lib().pfo_encode_blocks(ctx.h, C.byref(lay.s), nb, _p(x, i32p), _p(off, i64p), _p(nv, u32p), _p(diag, u64p),
                            _p(norm, u64p), nthreads or max_threads())
    return diag, norm


def _key_ptrs(keys):
    arr = (u64p * max(1, len(keys)))()
    keep = []
    for i, k in enumerate(keys):
        kk = np.ascontiguousarray(k, dtype=np.uint64)
        keep.append(kk)
        arr[i] = _p(kk, u64p)
    return arr, keep


def rotate_query_set(ctx: Context, lay: LayoutPlan, cts: np.ndarray, keys, chain: bool) -> np.ndarray:
    """cts [m][2][L][n] coefficient form -> [K][2][L][n] NTT form"""
    c = np.ascontiguousarray(cts, dtype=np.uint64).reshape(lay.m, 2, ctx.L, ctx.n)
    rot = np.zeros((lay.K, 2, ctx.L, ctx.n), dtype=np.uint64)
    arr, keep = _key_ptrs(keys)
    lib().pfo_rotate_query_set(ctx.h, C.byref(lay.s), _p(c, u64p), arr, int(chain), _p(rot, u64p))
    return rot


def block_distance(ctx: Context, lay: LayoutPlan, rot: np.ndarray, diag: np.ndarray, norm: np.ndarray) -> np.ndarray:
    out = np.zeros((2, ctx.L, ctx.n), dtype=np.uint64)
    lib().pfo_block_distance(ctx.h, C.byref(lay.s), _p(np.ascontiguousarray(rot), u64p),
                             _p(np.ascontiguousarray(diag), u64p), _p(np.ascontiguousarray(norm), u64p), _p(out, u64p))
    return out


def max_threads() -> int:
    """OpenMP's default team size (follows OMP_NUM_THREADS, which torchrun sets to 1)"""
    return min(lib().pfo_max_threads(), os.cpu_count() or 1)


def host_cores() -> int:
    """cores this process may run on — the thread count the CPU baseline uses (passed explicitly to the
    num_threads clauses, so an inherited OMP_NUM_THREADS=1 does not shrink it)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def search_pairs(ctx: Context, lay: LayoutPlan, cts: np.ndarray, keys, chain: bool, pair_query, pair_block,
                 diag: np.ndarray, norm: np.ndarray, nthreads: int = 1, result_limbs: int = 0):
    """Whole step on CPU.  cts [nq][m][2][L][n]; returns (out [P][2][Lr][n], (rot_s, mac_s)); result_limbs
    in [1, L) mod-switches every result down (SEAL mod_switch_to_inplace) as the CUDA engine does."""
    nq = cts.shape[0]
    pq = np.ascontiguousarray(pair_query, dtype=np.int32)
    pb = np.ascontiguousarray(pair_block, dtype=np.int64)
    P = len(pq)
    rot = np.zeros((nq, lay.K, 2, ctx.L, ctx.n), dtype=np.uint64)
    times = (C.c_double * 2)()
    arr, keep = _key_ptrs(keys)
    if result_limbs and result_limbs < ctx.L:
        out = np.zeros((P, 2, result_limbs, ctx.n), dtype=np.uint64)
        lib().pfo_search_pairs_ms(ctx.h, C.byref(lay.s), nq, _p(np.ascontiguousarray(cts, dtype=np.uint64), u64p), arr,
                                  int(chain), P, _p(pq, i32p), _p(pb, i64p), _p(np.ascontiguousarray(diag), u64p),
                                  _p(np.ascontiguousarray(norm), u64p), _p(rot, u64p), result_limbs, _p(out, u64p),
                                  nthreads, times)
        return out, (times[0], times[1])
    out = np.zeros((P, 2, ctx.L, ctx.n), dtype=np.uint64)
    lib().pfo_search_pairs(ctx.h, C.byref(lay.s), nq, _p(np.ascontiguousarray(cts, dtype=np.uint64), u64p), arr,
                           int(chain), P, _p(pq, i32p), _p(pb, i64p), _p(np.ascontiguousarray(diag), u64p),
                           _p(np.ascontiguousarray(norm), u64p), _p(rot, u64p), _p(out, u64p), nthreads, times)
    return out, (times[0], times[1])


# ---- plaintext path ----

def coarse_quantize(x: np.ndarray, centroids: np.ndarray, nprobe: int):
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    nq, d = x.shape
    idx = np.zeros((nq, nprobe), dtype=np.int64)
    dist = np.zeros((nq, nprobe), dtype=np.float32)
    lib().pfo_coarse_quantize(nq, d, c.shape[0], _p(x, f32p), _p(c, f32p), nprobe, _p(idx, i64p), _p(dist, f32p))
    return idx, dist


def search_lists_plain(x, idx, list_offsets, ids, vectors):
    x = np.ascontiguousarray(x, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    lo = np.ascontiguousarray(list_offsets, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    nq, d = x.shape
    nprobe = idx.shape[1]
    sizes = (lo[1:] - lo[:-1])
    cap = int(sum(int(sizes[l]) for l in idx.reshape(-1) if l >= 0))
    dist = np.zeros(max(cap, 1), dtype=np.float32)
    labels = np.zeros(max(cap, 1), dtype=np.int64)
    ls = (C.c_size_t * nq)()
    w = lib().pfo_search_lists_plain(nq, d, _p(x, f32p), _p(idx, i64p), nprobe, _p(lo, i64p), _p(ids, i64p),
                                     _p(v, f32p), _p(dist, f32p), _p(labels, i64p), cap, ls)
    assert w == cap
    return dist[:cap], labels[:cap], np.array(list(ls), dtype=np.int64)


def search_lists_pq(x, idx, centroids, list_offsets, ids, pq_M, pq_centroids, codes):
    """PQ-ADC restatement of the FAISS fork's search_encrypted (pf_oracle.c: pfo_search_lists_pq)"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    cent = np.ascontiguousarray(centroids, dtype=np.float32)
    lo = np.ascontiguousarray(list_offsets, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    pqc = np.ascontiguousarray(pq_centroids, dtype=np.float32).reshape(-1)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    nq, d = x.shape
    assert d % pq_M == 0 and pqc.size == 256 * d and codes.shape == (int(lo[-1]), pq_M)
    sizes = (lo[1:] - lo[:-1])
    cap = int(sum(int(sizes[l]) for l in idx.reshape(-1) if l >= 0))
    dist = np.zeros(max(cap, 1), dtype=np.float32)
    labels = np.zeros(max(cap, 1), dtype=np.int64)
    ls = (C.c_size_t * nq)()
    w = lib().pfo_search_lists_pq(nq, d, _p(x, f32p), _p(idx, i64p), idx.shape[1], _p(cent, f32p), _p(lo, i64p), _p(ids, i64p), pq_M,
                                  _p(pqc, f32p), codes.ctypes.data_as(C.POINTER(C.c_uint8)), _p(dist, f32p), _p(labels, i64p), cap, ls)
    assert w == cap
    return dist[:cap], labels[:cap], np.array(list(ls), dtype=np.int64)


def pq_encode_residuals(vectors, list_offsets, centroids, pq_M, pq_centroids):
    """codes [ntotal][M] of list-ordered vectors (FAISS compute_code on the residual to the list's centroid)"""
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    lo = np.ascontiguousarray(list_offsets, dtype=np.int64)
    cent = np.ascontiguousarray(centroids, dtype=np.float32)
    pqc = np.ascontiguousarray(pq_centroids, dtype=np.float32).reshape(-1)
    list_of = np.repeat(np.arange(len(lo) - 1, dtype=np.int64), np.diff(lo))
    codes = np.zeros((len(v), pq_M), dtype=np.uint8)
    lib().pfo_pq_encode_residuals(len(v), v.shape[1], _p(v, f32p), _p(list_of, i64p), _p(cent, f32p), pq_M, _p(pqc, f32p),
                                  codes.ctypes.data_as(C.POINTER(C.c_uint8)))
    return codes


def recall(returned: np.ndarray, gt: np.ndarray):
    r = np.ascontiguousarray(returned, dtype=np.int64)
    g = np.ascontiguousarray(gt, dtype=np.int32)
    outs = [C.c_double() for _ in range(5)]
    lib().pfo_recall(r.shape[0], r.shape[1], _p(r, i64p), g.shape[1], _p(g, i32p), *[C.byref(o) for o in outs])
    keys = ["ref_recall_1", "ref_recall_10", "ref_recall_100", "std_recall_10", "mrr_10"]
    return {k: o.value for k, o in zip(keys, outs)}


def vecs_read(path: str, dtype=np.float32) -> np.ndarray:
    d, n, data = C.c_size_t(), C.c_size_t(), C.c_void_p()
    rc = lib().pfo_vecs_read(path.encode(), C.byref(d), C.byref(n), C.byref(data))
    if rc:
        raise IOError(f"pfo_vecs_read({path}) -> {rc}")
    arr = np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint32)), shape=(n.value * d.value,)).copy()
    C.CDLL(None).free(data)
    return arr.view(dtype).reshape(n.value, d.value)


def blake2b_param(param: bytes, key: bytes, msg: bytes, outlen: int) -> bytes:
    pb, kb, mb = (np.frombuffer(x, dtype=np.uint8).copy() if len(x) else np.zeros(1, dtype=np.uint8) for x in (param, key, msg))
    out = np.zeros(outlen, dtype=np.uint8)
    u8 = C.POINTER(C.c_uint8)
    lib().pfo_blake2b_param(_p(pb, u8), _p(kb, u8), len(key), _p(mb, u8), len(msg), _p(out, u8), outlen)
    return out.tobytes()


def blake2xb(outlen: int, msg: bytes, key: bytes) -> bytes:
    kb, mb = (np.frombuffer(x, dtype=np.uint8).copy() if len(x) else np.zeros(1, dtype=np.uint8) for x in (key, msg))
    out = np.zeros(outlen, dtype=np.uint8)
    u8 = C.POINTER(C.c_uint8)
    lib().pfo_blake2xb(_p(out, u8), outlen, _p(mb, u8), len(msg), _p(kb, u8), len(key))
    return out.tobytes()
