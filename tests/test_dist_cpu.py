"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over the gloo backend: the ownership
rule + all-to-all routing of allpathslg_b200.dist, with the oracle standing in for the per-rank
sort/count kernels.  Identical k-mers must meet on one rank, so summed per-rank spectra equal the
single-process spectrum bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, out_dir, mode="hash"):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from allpathslg_b200 import _lib
    from allpathslg_b200.dist import host_partition_shuffle, host_shuffle
    from oracle import oracle_a as A

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _lib.lib()
    n_per = 3000
    sp = A.synth_params(40_000, 100)
    packed, off = A.synth_reads(sp, rank * n_per, n_per)  # this rank's slice of the reads
    # every canonical k-mer INSTANCE of this rank's reads, via the library's host-executable extractor
    W = (2 * K + 63) // 64
    total = int(off[-1])
    km = np.zeros(total * W, dtype=np.uint64)
    va = np.zeros(total, dtype=np.uint8)
    assert L.apgk_debug_host_extract(packed.ctypes.data, off.ctypes.data, n_per, K, km.ctypes.data, va.ctypes.data) == 0
    inst = km.reshape(-1, W)[va.astype(bool)]
    rng = (0, 0)
    if mode == "hash":
        mine = host_shuffle(inst, K, rank, world)
    else:
        mine, rng = host_partition_shuffle(inst, K, 12, rank, world)
    # sort + count the shard this rank owns
    if len(mine):
        order = np.lexsort(tuple(mine[:, j] for j in range(W - 1, -1, -1)))
        srt = mine[order]
        new = np.ones(len(srt), dtype=bool)
        new[1:] = (srt[1:] != srt[:-1]).any(axis=1)
        idx = np.nonzero(new)[0]
        counts = np.diff(np.append(idx, len(srt)))
        kmers = srt[idx]
    else:
        counts = np.zeros(0, dtype=np.int64)
        kmers = np.zeros((0, W), dtype=np.uint64)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), kmers=kmers, counts=counts, rng=np.array(rng))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,K", [(2, 25), (3, 40), (2, 96)])
def test_hash_sharded_count_gloo(oracle, tmp_path, world, K):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, K, str(tmp_path)), nprocs=world, join=True)
    sp = oracle.synth_params(40_000, 100)
    packed, off = oracle.synth_reads(sp, 0, 3000 * world)
    ek, ec, en = oracle.count(packed, off, K)
    from allpathslg_b200 import owner_of

    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # shards are disjoint, each holds exactly the k-mers it owns, counts are final
    own = owner_of(K, ek, world)
    for r, pt in enumerate(parts):
        sel = own == r
        assert (pt["kmers"] == ek[sel]).all()
        assert (pt["counts"].astype(np.uint64) == ec[sel]).all()
    spec = np.zeros(int(ec.max()) + 1, dtype=np.uint64)
    for pt in parts:
        spec += np.bincount(pt["counts"], minlength=len(spec)).astype(np.uint64)
    assert (spec == oracle.spectrum(ec)).all()
    assert sum(int(pt["counts"].sum()) for pt in parts) == en


@pytest.mark.parametrize("world,K", [(2, 25), (3, 40), (2, 96)])
def test_partition_first_sharded_count_gloo(oracle, tmp_path, world, K):
    """Host mirror of dist.sharded_count's exchange: bucket ranges of the 12-bit prefix, cut by
    balanced_splitters from the all-gathered histogram, routed with all_to_all over gloo."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, K, str(tmp_path), "partition"), nprocs=world, join=True)
    sp = oracle.synth_params(40_000, 100)
    packed, off = oracle.synth_reads(sp, 0, 3000 * world)
    ek, ec, en = oracle.count(packed, off, K)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    W = ek.shape[1]
    top = 2 * K - 64 * (W - 1)
    bucket = (ek[:, 0] >> np.uint64(top - 12)).astype(np.int64)
    lo_prev = 0
    for r, pt in enumerate(parts):
        lo, hi = int(pt["rng"][0]), int(pt["rng"][1])
        assert lo == lo_prev and hi >= lo
        lo_prev = hi
        sel = (bucket >= lo) & (bucket < hi)
        assert (pt["kmers"] == ek[sel]).all()
        assert (pt["counts"].astype(np.uint64) == ec[sel]).all()
    assert lo_prev == 1 << 12
    per = [int(pt["counts"].sum()) for pt in parts]
    assert sum(per) == en and max(per) < 1.2 * en / world  # balanced shards
    # union in rank order is the sorted table; summed spectra equal the single-process spectrum
    allk = np.concatenate([pt["kmers"] for pt in parts])
    assert (allk == ek).all()
    spec = np.zeros(int(ec.max()) + 1, dtype=np.uint64)
    for pt in parts:
        spec += np.bincount(pt["counts"], minlength=len(spec)).astype(np.uint64)
    assert (spec == oracle.spectrum(ec)).all()


def test_balanced_splitters():
    from allpathslg_b200.dist import balanced_splitters

    assert balanced_splitters(np.array([5, 5]), 2) == [0, 1, 2]
    assert balanced_splitters(np.array([10, 0, 0, 10]), 2) == [0, 3, 4]
    assert balanced_splitters(np.zeros(8, dtype=np.int64), 3) == [0, 8, 8, 8]
    assert balanced_splitters(np.array([7]), 4) == [0, 0, 0, 0, 1] or balanced_splitters(np.array([7]), 4)[-1] == 1
    rng = np.random.default_rng(5)
    t = rng.integers(0, 1000, size=4096)
    for world in (1, 2, 3, 8):
        b = balanced_splitters(t, world)
        assert b[0] == 0 and b[-1] == 4096 and len(b) == world + 1 and all(x <= y for x, y in zip(b, b[1:]))
        per = [int(t[b[r]:b[r + 1]].sum()) for r in range(world)]
        assert sum(per) == int(t.sum()) and max(per) - min(per) <= 2 * 1000


def test_plan_rounds_host_logic():
    """k-mer-space rounds of the sharded form: consecutive level-0 bucket ranges within the capacity, covering
    every bucket once, a bucket above the capacity alone in its range; deterministic."""
    from allpathslg_b200.dist import plan_rounds

    rng = np.random.default_rng(3)
    for cap in (1, 50, 1000, 10 ** 9):
        tot = rng.integers(0, 400, size=257)
        tot[100] = 5000
        r = plan_rounds(tot, cap)
        assert r[0][0] == 0 and r[-1][1] == len(tot)
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        for lo, hi in r:
            assert hi > lo
            if int(tot[lo:hi].sum()) > cap:   # only a single bucket above the capacity (after empty ones) may exceed it
                assert int((tot[lo:hi] > 0).sum()) == 1 and tot[hi - 1] > 0
        assert r == plan_rounds(tot.tolist(), cap)
    assert plan_rounds([0, 0, 0], 10) == [(0, 3)]
    assert plan_rounds([], 10) == [(0, 0)]


def _merge_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from allpathslg_b200.dist import merge_sparse_spectra

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [{1: 5, 70000: 1}, {1: 7, 2: 3, 70000: 2, 80000: 1}, {}][rank]
    spec, ni, nd = merge_sparse_spectra(mine, 100 * (rank + 1), 10 * (rank + 1), world, torch.device("cpu"))
    np.save(os.path.join(out_dir, "m%d.npy" % rank), np.concatenate([spec, [ni, nd]]).astype(np.uint64))
    dist.destroy_process_group()


def test_rounds_spectrum_merge_gloo(tmp_path):
    """The reduction that ends the sharded form's k-mer-space rounds: per-rank sparse spectra (frequencies beyond the
    dense 65536 included, one rank with nothing at all) and the totals, summed over three gloo ranks."""
    world = 3
    mp.spawn(_merge_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(os.path.join(str(tmp_path), "m%d.npy" % r)) for r in range(world)]
    for o in outs:
        assert (o == outs[0]).all()
    spec, ni, nd = outs[0][:-2], int(outs[0][-2]), int(outs[0][-1])
    assert len(spec) == 80001 and spec[1] == 12 and spec[2] == 3 and spec[70000] == 3 and spec[80000] == 1
    assert int(spec.sum()) == 19 and (ni, nd) == (600, 60)


def test_device_splitter_rule_on_the_host(apgk_lib):
    """The ownership rule of the sharded form (k_total_sizes -> scan -> k_splitters), run on the host through the
    library's test hook: with no per-bucket charge it is dist.balanced_splitters; with one, the ranges have equal COST
    up to one bucket, stay contiguous and cover the bucket space, whatever the skew."""
    from allpathslg_b200.dist import balanced_splitters

    rnd = np.random.RandomState(5)
    for nb, world, skew in [(1 << 12, 8, "canonical"), (1 << 10, 3, "uniform"), (1 << 12, 5, "spiky"), (64, 8, "sparse"), (16, 4, "empty")]:
        if skew == "canonical":      # density 2(1 - x): what canonical k-mers look like over their leading bits
            sizes = (8000 * (1 - np.arange(nb) / nb) + rnd.randint(0, 50, nb)).astype(np.uint32)
        elif skew == "uniform":
            sizes = rnd.randint(3000, 5000, nb).astype(np.uint32)
        elif skew == "spiky":
            sizes = rnd.randint(0, 100, nb).astype(np.uint32)
            sizes[rnd.randint(0, nb, 20)] = 500_000
        elif skew == "sparse":
            sizes = np.zeros(nb, dtype=np.uint32)
            sizes[[3, 40]] = [10, 7]
        else:
            sizes = np.zeros(nb, dtype=np.uint32)
        for cost in (0, 800, 5000):
            bounds = np.zeros(world + 1, dtype=np.uint32)
            assert apgk_lib.apgk_debug_host_splitters(sizes.ctypes.data, nb, world, cost, bounds.ctypes.data) == 0
            b = bounds.astype(np.int64)
            assert b[0] == 0 and b[-1] == nb and (np.diff(b) >= 0).all()
            w = np.where(sizes > 0, sizes.astype(np.int64) + cost, 0)
            if cost == 0:
                assert b.tolist() == balanced_splitters(sizes.astype(np.int64), world)
            tot = int(w.sum())
            per = [int(w[b[r]:b[r + 1]].sum()) for r in range(world)]
            assert sum(per) == tot
            if tot:
                assert max(per) <= tot / world + int(w.max()) + 1     # equal cost up to one bucket
