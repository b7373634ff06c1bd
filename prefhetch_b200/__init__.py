"""prefhetch_b200 — B200-native engine for the server-side search hot path of PreFHEtch.

Host-side mirror of the reference's `Server` surface (ref: include/server/server_lib.h:25-49) over
the C ABI in include/prefhetch_b200.h.  All arithmetic runs in hand-written sm_100a CUDA kernels
(prefhetch_b200/csrc); nothing here falls back to the CPU or touches oracle/.
"""
from .engine import Engine, PfError, SearchResult, bfv_default_primes, batching_plain_modulus, parms_id, seal_stream_inflate, seal_ct_expand, seal_galois_keys_expand, seal_ct_expand_batch  # noqa: F401

__all__ = ["Engine", "PfError", "SearchResult", "bfv_default_primes", "batching_plain_modulus", "parms_id", "seal_stream_inflate", "seal_ct_expand", "seal_galois_keys_expand", "seal_ct_expand_batch"]
