"""GPU parity tests of the k-mer OCCURRENCE records (apgk_build_occurrences: the (read id, signed position)
payload of the reference's SortKmers records / KmerParcels batches), through the C ABI, against the CPU
oracle on the same inputs -- bit-exact run offsets, read ids and positions.  Run with `-m gpu` on a B200.
Parity is UNPINNED: the oracle is a spec-derived restatement, cross-checked by the
independent oracle B and hand-computed vectors in tests/golden."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))
RC = str.maketrans("ACGT", "TGCA")


def _counter(packed, off, K, uniform=None, **kw):
    from allpathslg_b200 import KmerCounter

    kc = KmerCounter(K, want_counts=True, **kw)
    if uniform:
        kc.add_reads_uniform(packed, uniform[0], uniform[1])
    else:
        kc.add_reads(packed, off)
    kc.finish()
    return kc


def _assert_occ_equal(oracle, kc, packed, off, K):
    ek, ec, en = oracle.count(packed, off, K)
    info = kc.build_occurrences()
    assert info["n_occ"] == en
    ero, erid, epos = oracle.occurrences(packed, off, K, ek, en)
    ro, rid, pos = kc.occurrences()
    gk, gc = kc.counts()
    assert (gk == ek).all() and (gc.astype(np.uint64) == ec).all()
    assert (ro == ero).all()
    assert (rid == erid).all()
    assert (pos == epos).all()
    return info


@pytest.mark.parametrize("v", GOLD["occ_hand"], ids=lambda v: "%s-K%d" % ("+".join(v["reads"]), v["K"]))
def test_occurrence_hand_vectors(oracle, v):
    p, o = oracle.pack_strings(v["reads"])
    kc = _counter(p, o, v["K"])
    kc.build_occurrences()
    ro, rid, pos = kc.occurrences()
    got = [[[int(a), int(b)] for a, b in zip(rid[int(ro[i]):int(ro[i + 1])], pos[int(ro[i]):int(ro[i + 1])])]
           for i in range(len(ro) - 1)]
    assert got == v["occ"]
    kc.close()


@pytest.mark.parametrize("s", GOLD["occ_synth"], ids=lambda s: "G%d-K%d" % (s["genome_len"], s["K"]))
def test_occurrence_synth_golden(oracle, s):
    sp = oracle.synth_params(s["genome_len"], s["read_len"])
    p, o = oracle.synth_reads(sp, 0, s["n_reads"])
    kc = _counter(p, o, s["K"], uniform=(s["n_reads"], s["read_len"]))
    kc.build_occurrences()
    ro, rid, pos = kc.occurrences()
    assert hashlib.sha256(ro.tobytes() + rid.tobytes() + pos.tobytes()).hexdigest() == s["occ_sha256"]
    kc.close()


@pytest.mark.parametrize("K", [1, 2, 3, 11, 16, 24, 25, 27, 31, 32, 33, 48, 64, 65, 96])
def test_occurrences_ragged_reads(oracle, K):
    """Every key width; reads shorter than K, empty reads (they keep their ids), duplicates, a reverse
    complement pair, low-complexity reads; sub-ranges of the table."""
    rnd = random.Random(2000 + K)
    reads = ["".join(rnd.choice("ACGT") for _ in range(rnd.choice([0, 1, K - 1, K, K + 1, K + 7, 150, 100])))
             for _ in range(300)]
    reads += ["", "A" * (K + 300), "ACGT" * 100, reads[3], "T" * (K + 20), "", "", "CG" * (K + 5), ""]
    reads.append(reads[7].translate(RC)[::-1])
    p, o = oracle.pack_strings(reads)
    kc = _counter(p, o, K)
    _assert_occ_equal(oracle, kc, p, o, K)
    # a sub-range of k-mers gives the matching slice
    ro, rid, pos = kc.occurrences()
    nd = len(ro) - 1
    if nd >= 3:
        a, n = nd // 3, nd // 2
        sro, srid, spos = kc.occurrences(a, n)
        assert (sro == ro[a:a + n + 1] - ro[a]).all()
        assert (srid == rid[int(ro[a]):int(ro[a + n])]).all() and (spos == pos[int(ro[a]):int(ro[a + n])]).all()
    kc.close()


@pytest.mark.parametrize("K", [5, 25, 40, 96])
def test_occurrences_long_runs(oracle, K):
    """Runs beyond one thread's insertion sort: a duplicated read (runs of 300: shared-memory network),
    poly-A and tandem repeats (runs of tens of thousands: the in-place network)."""
    rnd = random.Random(K)
    dup = "".join(rnd.choice("ACGT") for _ in range(K + 40))
    reads = ["A" * 200] * 400 + [dup] * 300 + ["ACGTACGTAC" * 20] * 150 + ["".join(rnd.choice("AC") for _ in range(120)) for _ in range(200)]
    rnd.shuffle(reads)
    p, o = oracle.pack_strings(reads)
    kc = _counter(p, o, K)
    info = _assert_occ_equal(oracle, kc, p, o, K)
    assert info["n_big_runs"] > 0
    kc.close()


@pytest.mark.parametrize("K,prefix_bits", [(16, 2), (13, 4), (25, 6), (31, 8)])
def test_occurrences_buckets_beyond_the_shared_table(oracle, K, prefix_bits):
    """Few prefix bits: thousands of distinct k-mers per bucket, more than k_occ_place keeps in shared memory
    (those buckets search the table in global memory); K=31 uses full-key elements."""
    sp = oracle.synth_params(60_000, 100)
    p, o = oracle.synth_reads(sp, 0, 12_000)
    kc = _counter(p, o, K, uniform=(12_000, 100), prefix_bits=prefix_bits)
    _assert_occ_equal(oracle, kc, p, o, K)
    kc.close()


@pytest.mark.parametrize("K,L,n", [(25, 100, 400_000), (20, 250, 60_000), (48, 150, 100_000), (96, 250, 50_000)])
def test_occurrences_synthetic_coverage(oracle, K, L, n):
    """BASELINE configs scaled to what the single-threaded oracle walks in seconds: 20x coverage with planted
    repeats and 0.5 % substitutions, uniform-length ingest."""
    sp = oracle.synth_params(n * L // 20, L)
    p, o = oracle.synth_reads(sp, 0, n)
    kc = _counter(p, o, K, uniform=(n, L))
    _assert_occ_equal(oracle, kc, p, o, K)
    kc.close()


@pytest.mark.parametrize("K", [25, 48])
def test_occurrences_after_kmer_space_rounds_and_batches(oracle, K):
    """The table built in several k-mer-space rounds from reads added in unaligned batches serves the
    same occurrences."""
    from allpathslg_b200 import KmerCounter

    sp = oracle.synth_params(300_000, 150)
    p, o = oracle.synth_reads(sp, 0, 15_000)
    ek, ec, en = oracle.count(p, o, K)
    kc = KmerCounter(K, max_round_keys=en // 3 + 1)
    for a, b in ((0, 4001), (4001, 4002), (4002, 15_000)):
        kc.add_reads(p, o[a:b + 1])
    kc.finish()
    assert kc.geometry()["n_rounds"] >= 2
    _assert_occ_equal(oracle, kc, p, o, K)
    # rebuilding after another finish gives the same records
    kc.finish()
    _assert_occ_equal(oracle, kc, p, o, K)
    kc.close()


def test_occurrence_properties_at_scale():
    """5 M reads (380 M instances): sampled ranges of the table are checked on the host -- every record's
    window has the k-mer as its canonical form on the strand its sign says, runs ascend by (read id,
    position), run lengths equal the counts -- and the totals are exact."""
    from allpathslg_b200 import KmerCounter, synth_params

    n, L, K = 5_000_000, 100, 25
    kc = KmerCounter(K)
    kc.synth_reads(synth_params(10_000_000, L), 0, n)
    kc.finish()
    ni, nd = kc.totals()
    info = kc.build_occurrences()
    assert info["n_occ"] == ni == n * (L - K + 1)
    packed = kc.export_reads()
    words = packed.view(np.uint64)
    mask = np.uint64((1 << (2 * K)) - 1)
    seen = 0
    for first in (0, nd // 2, nd - 200_000):
        k, c = kc.counts(first, 200_000)
        ro, rid, pos = kc.occurrences(first, 200_000)
        assert (np.diff(ro.astype(np.int64)) == c.astype(np.int64)).all()
        run = np.repeat(np.arange(200_000), c.astype(np.int64))
        key = rid.astype(np.int64) * 1000 + np.abs(pos)
        same = run[1:] == run[:-1]
        assert (key[1:][same] > key[:-1][same]).all()
        # rebuild each window's forward k-mer from the packed reads (base q at bits [2q, 2q+2), little-endian)
        q = rid.astype(np.uint64) * np.uint64(L) + (np.abs(pos) - 1).astype(np.uint64)
        fw = np.zeros(len(q), dtype=np.uint64)
        for j in range(K):
            b = (words[(q + np.uint64(j)) >> np.uint64(5)] >> (((q + np.uint64(j)) & np.uint64(31)) << np.uint64(1))) & np.uint64(3)
            fw = (fw << np.uint64(2)) | b
        rc = np.zeros(len(q), dtype=np.uint64)
        t = fw.copy()
        for j in range(K):
            rc = (rc << np.uint64(2)) | (np.uint64(3) - (t & np.uint64(3)))
            t >>= np.uint64(2)
        rc &= mask
        canon = np.minimum(fw, rc)
        assert (canon == k[run, 0]).all()
        assert ((pos < 0) == (rc < fw)).all()
        seen += len(q)
    assert seen > 0
    kc.close()


def test_occurrences_state_errors(oracle):
    from allpathslg_b200 import ApgkError, KmerCounter, _lib

    p, o = oracle.pack_strings(["ACGTACGTAGCTAGCTAGCTAGGATCGATCGATTTAGC"])
    kc = KmerCounter(5, want_counts=True)
    kc.add_reads(p, o)
    with pytest.raises(ApgkError) as e:
        kc.build_occurrences()          # before finish
    assert e.value.code == _lib.E_STATE
    kc.finish()
    with pytest.raises(ApgkError) as e:
        kc.occurrences()                # before build
    assert e.value.code == _lib.E_STATE
    kc.build_occurrences()
    ro, _, _ = kc.occurrences()
    with pytest.raises(ApgkError) as e:
        kc.occurrences(len(ro), 5)      # beyond the table
    assert e.value.code == _lib.E_ARG
    kc.add_reads(p, o)                  # new reads invalidate the records
    with pytest.raises(ApgkError) as e:
        kc.occurrences()
    assert e.value.code == _lib.E_STATE
    kc.close()
    kc = KmerCounter(5, want_counts=False)
    kc.add_reads(p, o)
    kc.finish()
    with pytest.raises(ApgkError) as e:
        kc.build_occurrences()          # no table was requested
    assert e.value.code == _lib.E_STATE
    kc.close()


def test_reference_named_record_entry_points(oracle):
    """SortKmers(records=True) = one record per instance; KmerParcelsBuilder.Batches = k-mer + its list."""
    from allpathslg_b200 import KmerParcelsBuilder, SortKmers

    sp = oracle.synth_params(50_000, 100)
    p, o = oracle.synth_reads(sp, 0, 8_000)
    ek, ec, en = oracle.count(p, o, 25)
    ero, erid, epos = oracle.occurrences(p, o, 25, ek, en)
    k, rid, pos = SortKmers(p, o, 25, records=True)
    assert len(k) == en and (k == np.repeat(ek, ec.astype(np.int64), axis=0)).all()
    assert (rid == erid).all() and (pos == epos).all()
    b = KmerParcelsBuilder(25, p, o).Build()
    bk, bro, brid, bpos = b.Batches(100, 1000)
    assert (bk == ek[100:1100]).all() and (bro == ero[100:1101] - ero[100]).all()
    assert (brid == erid[int(ero[100]):int(ero[1100])]).all() and (bpos == epos[int(ero[100]):int(ero[1100])]).all()
    b.close()


def test_direct_sweep_agrees_with_two_phase_build(oracle, tmp_path):
    """APGK_OCC_DIRECT=1 selects the first implementation (lookup + slot + store straight from the sweep of
    the reads); it is read at library load, so it runs in a child process.  Both device paths must give the
    same records as the oracle."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, hashlib; sys.path.insert(0, %r)\n"
        "from oracle import oracle_a as A\n"
        "from allpathslg_b200 import KmerCounter\n"
        "for K, L, n in ((25, 100, 30000), (40, 120, 8000)):\n"
        "    p, o = A.synth_reads(A.synth_params(n * L // 20, L), 0, n)\n"
        "    kc = KmerCounter(K); kc.add_reads(p, o); kc.finish(); info = kc.build_occurrences()\n"
        "    ro, rid, pos = kc.occurrences()\n"
        "    print(K, info['ms']['place'] == 0.0, hashlib.sha256(ro.tobytes() + rid.tobytes() + pos.tobytes()).hexdigest())\n"
    ) % root
    outs = {}
    for mode in ("direct", "two-phase", "two-phase-unpacked"):
        env = dict(os.environ)
        env.pop("APGK_OCC_DIRECT", None)
        env.pop("APGK_OCC_NOPACK", None)
        if mode == "direct":
            env["APGK_OCC_DIRECT"] = "1"
        if mode == "two-phase-unpacked":   # remainder and position in two arrays (what > 2^31 bases with REM = 32 use)
            env["APGK_OCC_NOPACK"] = "1"
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr
        outs[mode] = [ln.split() for ln in r.stdout.strip().splitlines()]
    assert [x[1] for x in outs["direct"]] == ["True", "True"] and [x[1] for x in outs["two-phase"]] == ["False", "False"]
    assert [x[2] for x in outs["direct"]] == [x[2] for x in outs["two-phase"]] == [x[2] for x in outs["two-phase-unpacked"]]
    for (K, L, n), row in zip(((25, 100, 30000), (40, 120, 8000)), outs["two-phase"]):
        p, o = oracle.synth_reads(oracle.synth_params(n * L // 20, L), 0, n)
        ek, ec, en = oracle.count(p, o, K)
        ero, erid, epos = oracle.occurrences(p, o, K, ek, en)
        assert row[2] == hashlib.sha256(ero.tobytes() + erid.tobytes() + epos.tobytes()).hexdigest()


@pytest.mark.parametrize("K,prefix_bits", [(25, 0), (24, 0), (16, 4), (31, 0), (48, 0), (96, 0)])
def test_bulk_read_freqs_on_device(oracle, K, prefix_bits):
    """apgk_read_freqs_device: the frequency of every window of the store, resolved bucket by bucket, into a
    device buffer -- against the oracle's per-position frequencies and against the per-window table search."""
    import torch

    sp = oracle.synth_params(80_000, 100)
    p, o = oracle.synth_reads(sp, 0, 16_000)
    kc = _counter(p, o, K, uniform=(16_000, 100), prefix_bits=prefix_bits)
    ek, ec, en = oracle.count(p, o, K)
    erf = oracle.read_freqs(p, o, K, ek, ec)
    erf32 = np.where(erf == np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0xFFFFFFFF), erf).astype(np.uint32)
    tb, _ = kc.read_store_info()
    out = torch.zeros(tb, dtype=torch.int32, device="cuda")
    ms = kc.read_freqs_device(out.data_ptr())
    assert ms["sweep"] > 0 and ms["place"] > 0
    got = out.cpu().numpy().view(np.uint32)
    assert (got == erf32).all()
    half = tb // 2
    direct = np.concatenate([kc.read_freqs(0, half), kc.read_freqs(half, tb - half)])
    assert (direct == erf32).all()
    # the occurrence records survive a bulk lookup in between (they share the scatter buffers, not the result)
    kc.build_occurrences()
    a = kc.occurrences()
    kc.read_freqs_device(out.data_ptr())
    b = kc.occurrences()
    assert all((x == y).all() for x, y in zip(a, b))
    kc.close()


def test_no_instances_at_all(oracle):
    """Reads exist but none reaches K bases: an empty table, empty records, every window invalid."""
    from allpathslg_b200 import KmerCounter

    p, o = oracle.pack_strings(["ACGT", "", "ACGTACGTAC", "TT"])
    kc = KmerCounter(25, want_counts=True)
    kc.add_reads(p, o)
    kc.finish()
    assert kc.totals() == (0, 0)
    info = kc.build_occurrences()
    assert info["n_occ"] == 0
    ro, rid, pos = kc.occurrences()
    assert list(ro) == [0] and len(rid) == 0 and len(pos) == 0
    rf = kc.read_freqs()
    assert len(rf) == 16 and (rf == 0xFFFFFFFF).all()
    kc.close()
