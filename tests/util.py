"""Shared helpers for tests (synthetic data, toy parameters)."""
import numpy as np


def is_prime(n: int) -> bool:
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def ntt_primes(n: int, bits: int, count: int):
    """`count` largest `bits`-bit primes = 1 mod 2n"""
    out, p = [], ((1 << bits) - 1) // (2 * n) * (2 * n) + 1
    while len(out) < count:
        if p < (1 << bits) and is_prime(p):
            out.append(p)
        p -= 2 * n
    return out


def toy_params(n: int = 1024, bits: int = 40, k: int = 4, tbits: int = 20):
    primes = ntt_primes(n, bits, k - 1) + ntt_primes(n, bits + 1, 1)
    t = ntt_primes(n, tbits, 1)[0]
    return n, primes, t


def sift_like(rng: np.random.Generator, nb: int, d: int, nlist: int, nq: int):
    """SIFT-shaped synthetic data: uint8-valued float vectors from a Gaussian mixture."""
    centres = rng.uniform(0, 160, size=(nlist, d))
    assign = rng.integers(0, nlist, size=nb)
    base = np.clip(np.rint(centres[assign] + rng.normal(0, 24, size=(nb, d))), 0, 255).astype(np.float32)
    qa = rng.integers(0, nlist, size=nq)
    query = np.clip(np.rint(centres[qa] + rng.normal(0, 24, size=(nq, d))), 0, 255).astype(np.float32)
    return base, query, centres.astype(np.float32)


def build_ivf(base: np.ndarray, centroids: np.ndarray):
    """Assign every base vector to its nearest centroid; returns (list_offsets, ids, vectors in list order)."""
    d2 = ((base[:, None, :].astype(np.float64) - centroids[None, :, :].astype(np.float64)) ** 2).sum(-1) \
        if base.shape[0] * centroids.shape[0] <= 4_000_000 else None
    if d2 is None:
        b2 = (base.astype(np.float64) ** 2).sum(1)[:, None]
        c2 = (centroids.astype(np.float64) ** 2).sum(1)[None, :]
        d2 = b2 + c2 - 2.0 * base.astype(np.float64) @ centroids.astype(np.float64).T
    assign = d2.argmin(1)
    order = np.argsort(assign, kind="stable")
    counts = np.bincount(assign, minlength=centroids.shape[0])
    offsets = np.zeros(centroids.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    return offsets, order.astype(np.int64), np.ascontiguousarray(base[order])


class OracleClient:
    """Client side of the protocol (keygen / encrypt / decrypt), played by the CPU oracle in tests.
    A real deployment uses Microsoft SEAL here; the server never sees these secrets."""

    def __init__(self, oracle, n, primes, t, d, m, g, seed=2025):
        self.o = oracle
        self.ctx = oracle.Context(n, primes, t)
        self.lay = oracle.LayoutPlan(n, d, m, g)
        self.sk = self.ctx.keygen(seed)
        self.seed = seed
        self.keys = {}

    def galois_key(self, step):
        if step not in self.keys:
            self.keys[step] = self.ctx.galois_keygen(self.sk, self.ctx.galois_elt(step), self.seed + 1000 + step)
        return self.keys[step]

    def step_keys(self):
        return [self.galois_key(r) for r in range(1, self.lay.R)]

    def encrypt_query(self, q, seed):
        """-> [m][2][L][n] coefficient form"""
        qi = np.asarray(q).astype(np.int64)
        return np.stack([self.ctx.encrypt(self.sk, self.ctx.encode(self.lay.query_slots(self.ctx.t, qi, a)), seed + a)
                         for a in range(self.lay.m)])

    def serialize_queries(self, cts):
        """cts [nq][m][2][L][n] -> (blob uint8, offsets)"""
        blobs = [self.ctx.ct_save(ct) for q in cts for ct in q]
        offs = np.zeros(len(blobs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in blobs])
        return np.frombuffer(b"".join(blobs), dtype=np.uint8).copy(), offs

    def mod_switch_to(self, ct, limbs):
        """SEAL Evaluator::mod_switch_to_inplace down to `limbs` data limbs (oracle)"""
        while ct.shape[1] > limbs:
            ct = self.ctx.mod_switch_next(ct)
        return ct

    def distances(self, result_ct, q, nvec):
        limbs = result_ct.shape[1]
        if limbs < self.ctx.L:   # decrypt at the lower level: context / secret key restricted to its primes
            key = (limbs,)
            if not hasattr(self, "_low"):
                self._low = {}
            if key not in self._low:
                pr = self.ctx.primes[:limbs] + [self.ctx.primes[-1]]
                self._low[key] = (self.o.Context(self.ctx.n, pr, self.ctx.t),
                                  np.ascontiguousarray(self.sk[[*range(limbs), self.ctx.k - 1]]))
            ctx, sk = self._low[key]
            plain, budget = ctx.decrypt(sk, result_ct)
            slots = ctx.decode(plain).astype(np.int64)
            tab = self.lay.slot_table()
            qi = np.asarray(q).astype(np.int64)
            d = (slots[tab].sum(axis=1) + int((qi * qi).sum())) % self.ctx.t
            return d[:nvec], budget
        plain, budget = self.ctx.decrypt(self.sk, result_ct)
        slots = self.ctx.decode(plain).astype(np.int64)
        tab = self.lay.slot_table()
        qi = np.asarray(q).astype(np.int64)
        d = (slots[tab].sum(axis=1) + int((qi * qi).sum())) % self.ctx.t
        return d[:nvec], budget


def galois_keys_save(ctx, keys_by_elt: dict, parms_id=(0, 0, 0, 0)) -> bytes:
    """SEAL 4.1 GaloisKeys::save with compr_mode none ([EXT] kswitchkeys.h KSwitchKeys::save_members inside a
    Serialization::Save envelope): SEALHeader, parms_id, dim1 = N key slots (GaloisKeys resizes to
    coeff_count), per slot dim2 and dim2 PublicKey streams — each a SEAL envelope around the members of a
    size-2 NTT-form ciphertext over all k primes.  Slot of element e is (e - 1) / 2.
    keys_by_elt: {galois_elt: words [L][2][k][n]}."""
    import struct
    n, k, L = ctx.n, ctx.k, ctx.L
    parts = [struct.pack("<4Q", *parms_id), struct.pack("<Q", n)]
    slots = {(e - 1) // 2: w for e, w in keys_by_elt.items()}
    for i in range(n):
        if i not in slots:
            parts.append(struct.pack("<Q", 0))
            continue
        w = np.asarray(slots[i], dtype=np.uint64).reshape(L, 2, k, n)
        parts.append(struct.pack("<Q", L))
        for j in range(L):
            parts.append(ctx.ct_save(w[j], is_ntt=True, parms_id=parms_id))
    body = b"".join(parts)
    total = 16 + len(body)
    return bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0]) + struct.pack("<Q", total) + body


def zlib_stream(raw: bytes) -> bytes:
    """what SEAL writes with compr_mode_type::zlib: header (compr_mode 1, new size) + deflate of the body"""
    import struct
    import zlib
    body = zlib.compress(raw[16:], 6)
    return raw[:5] + b"\x01" + raw[6:8] + struct.pack("<Q", 16 + len(body)) + body


def have_zstd() -> bool:
    try:
        import pyarrow as pa
        return bool(pa.Codec.is_available("zstd"))
    except Exception:       # noqa: BLE001
        return False


def zstd_stream(raw: bytes, streaming: bool = False) -> bytes:
    """what SEAL writes with compr_mode_type::zstd: header (compr_mode 2, new size) + one Zstandard frame of the
    body.  The frames come from pyarrow's bundled libzstd: one-shot (content size in the frame header) or, like
    SEAL's ZSTD_compressStream2 loop, streamed (no content size)."""
    import struct
    import pyarrow as pa
    if streaming:
        sink = pa.BufferOutputStream()
        with pa.CompressedOutputStream(sink, "zstd") as z:
            z.write(raw[16:])
        body = sink.getvalue().to_pybytes()
    else:
        body = pa.compress(raw[16:], codec="zstd", asbytes=True)
    return raw[:5] + b"\x02" + raw[6:8] + struct.pack("<Q", 16 + len(body)) + body
