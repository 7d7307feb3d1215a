"""CPU tests of the library's host-executable pieces: the inline device helpers run on the host
through the apgk_debug_host_* hooks, the C ABI surface, and argument checking.  No GPU needed."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from oracle import oracle_b as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_extract(L, packed, off, K):
    W = (2 * K + 63) // 64
    total = int(off[-1] - off[0])
    km = np.zeros(max(total, 1) * W, dtype=np.uint64)
    va = np.zeros(max(total, 1), dtype=np.uint8)
    assert L.apgk_debug_host_extract(packed.ctypes.data, off.ctypes.data, len(off) - 1, K, km.ctypes.data,
                                     va.ctypes.data) == 0
    return km.reshape(-1, W)[:total], va[:total]


@pytest.mark.parametrize("K", [1, 2, 3, 8, 15, 16, 17, 24, 25, 31, 32, 33, 48, 63, 64, 65, 80, 95, 96])
def test_extract_helpers_match_oracle_b(apgk_lib, oracle, K):
    """extract16 / window_valid_mask16 (the code the kernels inline) against string-based oracle B,
    window by window, on ragged reads including reads shorter than K and empty reads."""
    rnd = random.Random(100 + K)
    reads = ["".join(rnd.choice("ACGT") for _ in range(rnd.choice([0, 1, 2, K - 1, K, K + 1, K + 7, 100, 33])))
             for _ in range(40)]
    reads += ["A" * (K + 30), "ACGT" * 30, "T" * K]
    p, o = oracle.pack_strings(reads)
    km, va = _host_extract(apgk_lib, p, o, K)
    W = (2 * K + 63) // 64
    exp_valid = np.zeros(int(o[-1]), dtype=np.uint8)
    for r, s in enumerate(reads):
        for q in range(len(s) - K + 1):
            pos = int(o[r]) + q
            exp_valid[pos] = 1
            want = B.kmer_to_int(B.canonical_str(s[q:q + K]))
            got = sum(int(km[pos, j]) << (64 * (W - 1 - j)) for j in range(W))
            assert got == want
    assert (exp_valid == va).all()


@pytest.mark.parametrize("K,D", [(25, 10), (25, 11), (25, 12), (20, 9), (33, 12), (64, 11), (96, 12), (8, 10), (6, 12), (13, 1)])
def test_top_digit_identity(apgk_lib, oracle, K, D):
    """top_digits16 (cheap level-0 histogram): top D bits of canonical == min(top D of fw, top D of rc),
    computed from the first / last ceil(D/2) bases only -- against the full extraction, window by window."""
    rnd = random.Random(K * 100 + D)
    reads = ["".join(rnd.choice("ACGT") for _ in range(rnd.choice([K, K + 1, K + 20, 130]))) for _ in range(40)]
    reads += ["A" * (K + 20), "ACGT" * 30, "T" * (K + 5), "AAAAAAAAAAAAAAAACCCCCCCCCCCCCCCCGGGGGGGGGGGGGGGGTTTTTTTTTTTTTTTT" * 3]
    p, o = oracle.pack_strings(reads)
    km, va = _host_extract(apgk_lib, p, o, K)
    W = (2 * K + 63) // 64
    dig = np.zeros(int(o[-1]), dtype=np.uint32)
    assert apgk_lib.apgk_debug_host_topdigits(p.ctypes.data, o.ctypes.data, len(o) - 1, K, D, dig.ctypes.data) == 0
    for pos in np.nonzero(va)[0]:
        full = sum(int(km[pos, j]) << (64 * (W - 1 - j)) for j in range(W))
        assert int(dig[pos]) == full >> (2 * K - D), (K, D, pos)


@pytest.mark.parametrize("K", [1, 5, 25, 32, 33, 64, 65, 96])
def test_canonical_helper(apgk_lib, K):
    rnd = random.Random(K)
    W = (2 * K + 63) // 64
    ks = ["".join(rnd.choice("ACGT") for _ in range(K)) for _ in range(100)] + ["A" * K, "T" * K, ("ACGT" * K)[:K]]
    arr = np.zeros((len(ks), W), dtype=np.uint64)
    for i, s in enumerate(ks):
        v = B.kmer_to_int(s)
        for j in range(W):
            arr[i, j] = (v >> (64 * (W - 1 - j))) & ((1 << 64) - 1)
    out = np.zeros_like(arr)
    assert apgk_lib.apgk_debug_host_canonical(K, arr.ctypes.data, len(ks), out.ctypes.data) == 0
    for i, s in enumerate(ks):
        want = B.kmer_to_int(B.canonical_str(s))
        assert sum(int(out[i, j]) << (64 * (W - 1 - j)) for j in range(W)) == want


def test_synth_generator_matches_oracle(apgk_lib, oracle):
    from allpathslg_b200 import synth_params

    for (G, L, r0, n) in [(100003, 100, 0, 500), (7001, 36, 17, 333), (250000, 250, 5, 200)]:
        so = oracle.synth_params(G, L)
        p, _ = oracle.synth_reads(so, r0, n)
        sl = synth_params(G, L)
        buf = np.zeros(len(p), dtype=np.uint8)
        assert apgk_lib.apgk_debug_host_synth(C.byref(sl), r0, n, buf.ctypes.data) == 0
        assert (buf == p).all()


def test_abi_exports_every_declared_symbol(apgk_lib):
    """Every function include/apgk.h declares is exported by libapgk.so and bound in _lib.SYMBOLS."""
    from allpathslg_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "apgk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(apgk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert getattr(apgk_lib, name) is not None


def test_words_per_kmer(apgk_lib):
    for K, W in [(1, 1), (25, 1), (32, 1), (33, 2), (64, 2), (65, 3), (96, 3)]:
        assert apgk_lib.apgk_words_per_kmer(K) == W


def test_owner_hash_is_balanced_and_stable(apgk_lib):
    from allpathslg_b200 import owner_of

    rnd = np.random.RandomState(0)
    for K, n_ranks in [(25, 8), (25, 2), (96, 4)]:
        W = (2 * K + 63) // 64
        q = rnd.randint(0, 2 ** 62, size=(20000, W), dtype=np.int64).astype(np.uint64)
        o1 = owner_of(K, q, n_ranks)
        o2 = owner_of(K, q, n_ranks)
        assert (o1 == o2).all() and o1.max() < n_ranks
        cnt = np.bincount(o1, minlength=n_ranks)
        assert cnt.min() > 0.8 * len(q) / n_ranks


def test_create_fails_loudly_without_gpu(apgk_lib):
    """No CPU fallback: on a box without a CUDA device apgk_create returns APGK_E_CUDA."""
    import torch

    from allpathslg_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = _lib.Config(K=25, device=0, flags=3, prefix_bits=0, reserve_bases=0, max_round_keys=0)
    h = C.c_void_p()
    assert apgk_lib.apgk_create(C.byref(cfg), C.byref(h)) == _lib.E_CUDA
    assert not h.value


def test_create_rejects_bad_k(apgk_lib):
    from allpathslg_b200 import _lib

    for K in (0, -3, 97):
        cfg = _lib.Config(K=K, device=0, flags=3, prefix_bits=0, reserve_bases=0, max_round_keys=0)
        h = C.c_void_p()
        assert apgk_lib.apgk_create(C.byref(cfg), C.byref(h)) == _lib.E_ARG


def test_kmer_spectrum_text_roundtrip(tmp_path):
    from allpathslg_b200 import KmerSpectrum

    s = KmerSpectrum(25, [0, 10, 3, 0, 0, 7, 20, 9])
    s.write(tmp_path / "x.kspec")
    t = KmerSpectrum.read(tmp_path / "x.kspec")
    assert t.K == 25 and (t.spec == s.spec).all()
    assert s.n_distinct() == 49 and s.n_instances() == 10 + 6 + 35 + 120 + 63
    est = s.estimate(read_len=100)
    assert est["kmer_coverage"] == 6.0


@pytest.mark.parametrize("K,P", [(25, 20), (25, 8), (13, 10), (5, 12), (3, 4), (31, 16), (32, 12), (33, 10), (48, 14), (64, 20), (96, 11)])
def test_table_search_on_host(apgk_lib, K, P):
    """table_find_index (prefix index + interpolated start + galloping bracket + binary search: the search
    every frequency-table lookup and the occurrence sweep run on the device) executed on the host, against
    a dict, for uniform and heavily skewed key sets, hits and misses."""
    rng = np.random.default_rng(K * 131 + P)
    W = (2 * K + 63) // 64
    bits = 2 * K
    for n in (0, 1, 5, 17, 40, 1000, 60000):
        if W == 1:
            keys = np.unique(rng.integers(0, 1 << min(bits, 63), size=n, dtype=np.uint64)).reshape(-1, 1)
        else:
            top = bits - 64 * (W - 1)
            cols = [rng.integers(0, 1 << min(top, 62), size=n, dtype=np.uint64)]
            cols += [rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
                     for _ in range(W - 1)]
            if n > 10:
                cols[0][: n // 2] = cols[0][0]   # half the keys in one prefix bucket, clustered
            keys = np.unique(np.stack(cols, 1), axis=0)
        m = len(keys)
        if m == 0:
            qs = np.zeros((3, W), dtype=np.uint64)
        else:
            miss = keys.copy()
            miss[:, -1] ^= np.uint64(1)
            qs = np.concatenate([keys, miss])
            if W == 1 and bits < 64:
                qs = qs[qs[:, 0] < np.uint64(1 << bits)]
        qs = np.ascontiguousarray(qs)
        keys = np.ascontiguousarray(keys)
        out = np.zeros(len(qs), dtype=np.uint64)
        assert apgk_lib.apgk_debug_host_table_find(K, keys.ctypes.data if m else None, m, min(P, 2 * K) if W > 1 else P,
                                                  qs.ctypes.data, len(qs), out.ctypes.data) == 0
        d = {tuple(k): i for i, k in enumerate(keys.tolist())}
        exp = np.array([d.get(tuple(x), 2 ** 64 - 1) for x in qs.tolist()], dtype=np.uint64)
        assert (out == exp).all(), (K, P, n)


def test_group_api_without_a_gpu(apgk_lib):
    """The group entry points answer without a device: NCCL is loaded on first use and gives an id (the bootstrap
    needs no GPU), null arguments are refused -- and nothing falls back to the CPU: a context cannot be made, so a
    group cannot be formed."""
    import ctypes as C

    import torch

    from allpathslg_b200 import _lib

    buf = np.zeros(128, dtype=np.uint8)
    rc = apgk_lib.apgk_group_unique_id(buf.ctypes.data)
    assert rc in (0, _lib.E_CUDA)          # E_CUDA only where libnccl.so.2 cannot be loaded at all
    if rc == 0:
        assert buf.any()
    h = C.c_void_p()
    assert apgk_lib.apgk_group_local(None, 0, C.byref(h)) == _lib.E_ARG
    assert apgk_lib.apgk_group_count(None) == _lib.E_ARG
    assert apgk_lib.apgk_group_totals(None, None, None) == _lib.E_ARG
    apgk_lib.apgk_group_destroy(None)      # a no-op
    if not torch.cuda.is_available():
        cfg = _lib.Config(K=25, device=0, flags=_lib.WANT_COUNTS)
        ctx = C.c_void_p()
        assert apgk_lib.apgk_create(C.byref(cfg), C.byref(ctx)) == _lib.E_CUDA   # no CPU fallback
