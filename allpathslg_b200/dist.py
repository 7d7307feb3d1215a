"""Multi-GPU k-mer spectrum: one process per GPU, sharded by canonical k-mer.

Partition-first form (`sharded_count`, the default).  Every rank runs levels 0 and 1 of the
single-GPU pipeline on ITS reads (`apgk_partition`): the canonical k-mers end up grouped by their
leading P bits, as 32-bit remainders when they fit.  The ranks all-gather the bucket histogram,
cut the bucket space into `world` contiguous ranges of (nearly) equal instance counts
(`balanced_splitters`), exchange whole ranges with one NCCL all-to-all over NVLink
(`torch.distributed.all_to_all_single`), and each rank sorts + counts the ranges it owns
(`apgk_count_pieces`).  The per-rank spectra -- disjoint sets of k-mers, so plain integer sums --
are all-reduced.  A k-mer's owner depends on the canonical k-mer alone, so counts are final
without a merge.

K-mer-space rounds (`_sharded_count_rounds`): when a rank's k-mers do not fit the device in one go
(`apgk_partition` answers APGK_E_RANGE), the ranks all-reduce their level-0 bucket totals, cut the
level-0 bucket space into the same consecutive ranges everywhere (`plan_rounds`) and run the
partition-first pipeline once per range (`apgk_partition_range`); a round's spectrum and totals are
harvested before the next round reuses the buffers, so the context keeps the LAST round's shard table
only (pass `on_round` to take every round's table).  Ranges ascend in k-mer order: the tables
concatenated in (round, rank) order are the globally sorted table.

Hash form (`sharded_count_hash`, the first implementation; APGK_SHARD_ROUNDS=0 selects it instead of the
rounds): owner = hash(k-mer) % world
(`apgk_owner_plan` / `apgk_owner_scatter`), exchange of full k-mers, then the whole pipeline again
on the received keys (`apgk_finish_keys_device`).  It extracts and partitions every k-mer twice
and sends 8 bytes per instance where the partition-first form sends 4.

torch is used for device buffers, streams and the collectives only.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def _splitters_tensor(bucket_totals, world):
    """bounds as an int64 tensor [world+1] on the device of bucket_totals (no host round trip)."""
    t = bucket_totals.to(torch.int64)
    nb = int(t.numel())
    dev = t.device
    if nb == 0:
        return torch.zeros(world + 1, dtype=torch.int64, device=dev)
    cum = torch.cumsum(t, 0)
    total = cum[-1]
    r = torch.arange(1, world, dtype=torch.int64, device=dev)
    targets = torch.div(total * r, world, rounding_mode="floor")
    inner = torch.searchsorted(cum, targets, right=True).clamp(max=nb)
    return torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), inner,
                      torch.full((1,), nb, dtype=torch.int64, device=dev)])


def balanced_splitters(bucket_totals, world):
    """Cut buckets [0, nb) into `world` contiguous ranges with (nearly) equal instance counts.
    bucket_totals: 1-D integer torch tensor or numpy array (instances per bucket, summed over ranks).
    -> list of world+1 bucket indices, bounds[0] = 0, bounds[world] = nb, non-decreasing: range r ends
    after the first bucket whose cumulative count exceeds total * (r+1) / world.
    Deterministic in its input, so every rank derives the same ownership."""
    if not isinstance(bucket_totals, torch.Tensor):
        bucket_totals = torch.from_numpy(np.ascontiguousarray(np.asarray(bucket_totals)).astype(np.int64))
    return [int(x) for x in _splitters_tensor(bucket_totals, world).cpu()]


def plan_rounds(level0_max, capacity):
    """Cut the level-0 buckets [0, len(level0_max)) into consecutive ranges whose summed entries stay within
    `capacity` (a single bucket above it gets a range of its own).  level0_max: per bucket, the largest count any
    rank holds.  Deterministic in its input, so every rank derives the same rounds.  -> list of (lo, hi)."""
    rounds, lo, acc = [], 0, 0
    for d, t in enumerate(int(x) for x in level0_max):
        if acc and acc + t > capacity:
            rounds.append((lo, d))
            lo, acc = d, 0
        acc += t
    rounds.append((lo, len(level0_max)))
    return rounds


def sharded_count(kc, rank, world, group=None, timings=None, on_round=None):
    """Run the sharded pipeline on this rank's KmerCounter `kc` (reads already in its store).

    Returns (spectrum uint64 array summed over all ranks, n_instances_global, n_distinct_global).
    The rank's own shard table stays queryable in `kc` (counts of the k-mers it owns); when the k-mers need
    several k-mer-space rounds that is the last round's shard -- `on_round(kc, i, n_rounds)` is called after
    every round for callers that want each round's table."""
    from .kmers import ApgkError

    dev = torch.device("cuda", torch.cuda.current_device())
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    e0.record()
    wall = {}
    t_last = [time.perf_counter()]

    def lap(name):  # host wall clock of the synchronous phases (diagnostics in `timings`)
        now = time.perf_counter()
        wall[name] = wall.get(name, 0.0) + 1e3 * (now - t_last[0])
        t_last[0] = now

    # ---- same geometry on every rank
    up = torch.tensor([kc.window_upper()], dtype=torch.int64, device=dev)
    dist.all_reduce(up, op=dist.ReduceOp.MAX, group=group)
    P = kc.choose_prefix_bits(int(up.item()))
    # ---- local partition (levels 0 + 1)
    lap("w_geometry")
    failed = 0
    try:
        kc.partition(P)
    except ApgkError as e:
        if e.code != -5:  # APGK_E_RANGE: more than one k-mer-space round needed here
            raise
        failed = 1
    lap("w_partition")
    sizes = None
    want_peer = world > 1 and os.environ.get("APGK_SHARD_EXCHANGE", "peer") == "peer"
    handle = np.zeros(64, dtype=np.uint8)
    if not failed:
        sizes_ptr, nb, elems_ptr, eb, n_elems = kc.partition_info()
        sizes = _wrap(sizes_ptr, nb, "<i8", dev)
        if want_peer:
            handle = kc.partition_export()
            d2, sub_ptr = kc.partition_subsizes(max(0, (world - 1).bit_length()))
    lap("w_subsizes")
    # one all-gather carries the bucket histogram, the "could not partition" flag and the IPC handle
    nb_all = 1 << P
    mine = torch.empty(nb_all + 9, dtype=torch.int64, device=dev)
    mine[:nb_all] = sizes if sizes is not None else 0
    mine[nb_all] = failed
    mine[nb_all + 1:] = torch.from_numpy(handle.view(np.int64).copy()).to(dev)
    gathered = torch.empty((world, nb_all + 9), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    all_sizes = gathered[:, :nb_all]
    bounds_t = _splitters_tensor(all_sizes.sum(0), world)
    cum = torch.zeros((world, nb_all + 1), dtype=torch.int64, device=dev)
    torch.cumsum(all_sizes, 1, out=cum[:, 1:])
    # everything the host needs in one transfer: flags, largest piece, bounds, cumulative counts at the bounds, handles
    small = torch.cat([gathered[:, nb_all].max().reshape(1), all_sizes.max().reshape(1), bounds_t,
                       cum[:, bounds_t].reshape(-1), gathered[:, nb_all + 1:].reshape(-1)]).cpu().numpy()
    lap("w_plan")
    if int(small[0]):
        if os.environ.get("APGK_SHARD_ROUNDS", "1") == "0":
            if timings is not None:
                timings["path"] = "hash"
            return sharded_count_hash(kc, rank, world, group, timings)
        return _sharded_count_rounds(kc, rank, world, P, dev, group, timings, on_round, want_peer)
    if int(small[1]) >= 2 ** 31:
        raise RuntimeError("a bucket piece holds 2^31 or more k-mers")
    bounds = [int(x) for x in small[2: 3 + world]]
    at_bounds = small[3 + world: 3 + world + world * (world + 1)].reshape(world, world + 1)
    handles = small[3 + world + world * (world + 1):].reshape(world, 8)
    lo, hi = bounds[rank], bounds[rank + 1]
    sizes_u32 = all_sizes.to(torch.int32).contiguous()
    # ---- exchange fused into the gather: map the peers' partition buffers (CUDA IPC over NVLink)
    peer_ptrs = None
    if want_peer:
        peer_ptrs = _open_peers(kc, rank, world, handles)
        ok = torch.tensor([1 if peer_ptrs is not None else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # every rank or none
        if not int(ok.item()):
            peer_ptrs = None
    if peer_ptrs is not None:
        e1.record()
        e2.record()
        torch.cuda.current_stream().synchronize()
        n_recv = int((at_bounds[:, rank + 1] - at_bounds[:, rank]).sum())
        # the senders' sub-bucket counts of my range (so that the pieces cross NVLink once)
        sub = _wrap(sub_ptr, nb_all << d2, "<i4", dev)
        sub_recv = torch.empty(world * ((hi - lo) << d2) + 1, dtype=torch.int32, device=dev)
        dist.all_to_all_single(sub_recv[: world * ((hi - lo) << d2)], sub,
                               output_split_sizes=[(hi - lo) << d2] * world,
                               input_split_sizes=[(bounds[r + 1] - bounds[r]) << d2 for r in range(world)], group=group)
        torch.cuda.current_stream().synchronize()
        lap("w_peer_setup")
        kc.count_pieces_peer(peer_ptrs, sizes_u32.data_ptr(), at_bounds[:, rank].astype(np.uint64), lo, hi,
                             split_bits=d2, d_sub_sizes=sub_recv.data_ptr())
        lap("w_count")
        e3.record()
    else:
        # ---- NCCL all-to-all into a receive buffer, then gather
        # per-destination send counts (my pieces) and per-source receive counts (their pieces of my range)
        send_counts = (at_bounds[rank, 1:] - at_bounds[rank, :-1]).astype(np.int64)
        recv_counts = (at_bounds[:, rank + 1] - at_bounds[:, rank]).astype(np.int64)
        n_recv = int(recv_counts.sum())
        words = 1 if eb == 4 else eb // 8
        tstr, tdt = ("<i4", torch.int32) if eb == 4 else ("<i8", torch.int64)
        send = _wrap(elems_ptr, max(n_elems, 1) * words, tstr, dev) if elems_ptr else torch.empty(1, dtype=tdt, device=dev)
        need = max(n_recv, 1) * eb
        buf = getattr(kc, "_recv_buf", None)
        if buf is None or buf.numel() * 8 < need:
            kc._recv_buf = None
            del buf
            torch.cuda.empty_cache()
            buf = torch.empty(int(need * 1.02) // 8 + 1024, dtype=torch.int64, device=dev)
            kc._recv_buf = buf
        recv = buf.view(tdt)
        e1.record()
        dist.all_to_all_single(recv[: n_recv * words], send[: n_elems * words],
                               output_split_sizes=[int(c) * words for c in recv_counts],
                               input_split_sizes=[int(c) * words for c in send_counts], group=group)
        e2.record()
        torch.cuda.current_stream().synchronize()
        del send
        seg_off = np.concatenate([[0], np.cumsum(recv_counts)[:-1]]).astype(np.uint64)
        lap("w_all_to_all")
        kc.count_pieces(recv.data_ptr(), world, sizes_u32.data_ptr(), seg_off, lo, hi)
        lap("w_count")
        e3.record()
    gather_ms = kc.stage_ms().get("owner", 0.0)
    if on_round is not None:
        on_round(kc, 0, 1)
    out = _reduce_results(kc, dev, group)
    lap("w_reduce")
    if timings is not None:
        torch.cuda.current_stream().synchronize()
        timings.update(wall)
        timings["gather_split_ms"] = gather_ms
        timings["path"] = "partition-first" + ("/peer" if peer_ptrs is not None else "/nccl")
        timings["partition_ms"] = e0.elapsed_time(e1)
        timings["all_to_all_ms"] = e1.elapsed_time(e2)
        timings["count_ms"] = e2.elapsed_time(e3)
        timings["sent_elems"] = int(n_elems)
        timings["recv_elems"] = n_recv
        timings["elem_bytes"] = int(eb)
        timings["bucket_range"] = (int(lo), int(hi))
        timings["prefix_bits"] = int(P)
    return out


def _exchange_and_count(kc, rank, world, P, dev, group, want_peer):
    """One partition-first exchange for the partition `kc` holds: all-gather of the bucket histograms, balanced
    ranges, the owned range gathered from the peers' buffers (or through an all-to-all) and counted.
    -> (bucket_lo, bucket_hi, path)."""
    sizes_ptr, nb, elems_ptr, eb, n_elems = kc.partition_info()
    sizes = _wrap(sizes_ptr, nb, "<i8", dev)
    handle = np.zeros(64, dtype=np.uint8)
    d2 = sub_ptr = None
    if want_peer:
        handle = kc.partition_export()
        d2, sub_ptr = kc.partition_subsizes(max(0, (world - 1).bit_length()))
    nb_all = 1 << P
    mine = torch.empty(nb_all + 8, dtype=torch.int64, device=dev)
    mine[:nb_all] = sizes
    mine[nb_all:] = torch.from_numpy(handle.view(np.int64).copy()).to(dev)
    gathered = torch.empty((world, nb_all + 8), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    all_sizes = gathered[:, :nb_all]
    bounds_t = _splitters_tensor(all_sizes.sum(0), world)
    cum = torch.zeros((world, nb_all + 1), dtype=torch.int64, device=dev)
    torch.cumsum(all_sizes, 1, out=cum[:, 1:])
    small = torch.cat([all_sizes.max().reshape(1), bounds_t, cum[:, bounds_t].reshape(-1),
                       gathered[:, nb_all:].reshape(-1)]).cpu().numpy()
    if int(small[0]) >= 2 ** 31:
        raise RuntimeError("a bucket piece holds 2^31 or more k-mers")
    bounds = [int(x) for x in small[1: 2 + world]]
    at_bounds = small[2 + world: 2 + world + world * (world + 1)].reshape(world, world + 1)
    handles = small[2 + world + world * (world + 1):].reshape(world, 8)
    lo, hi = bounds[rank], bounds[rank + 1]
    sizes_u32 = all_sizes.to(torch.int32).contiguous()
    peer_ptrs = None
    if want_peer:
        peer_ptrs = _open_peers(kc, rank, world, handles)
        ok = torch.tensor([1 if peer_ptrs is not None else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if not int(ok.item()):
            peer_ptrs = None
    if peer_ptrs is not None:
        sub = _wrap(sub_ptr, nb_all << d2, "<i4", dev)
        sub_recv = torch.empty(world * ((hi - lo) << d2) + 1, dtype=torch.int32, device=dev)
        dist.all_to_all_single(sub_recv[: world * ((hi - lo) << d2)], sub,
                               output_split_sizes=[(hi - lo) << d2] * world,
                               input_split_sizes=[(bounds[r + 1] - bounds[r]) << d2 for r in range(world)], group=group)
        torch.cuda.current_stream().synchronize()
        kc.count_pieces_peer(peer_ptrs, sizes_u32.data_ptr(), at_bounds[:, rank].astype(np.uint64), lo, hi,
                             split_bits=d2, d_sub_sizes=sub_recv.data_ptr())
        return lo, hi, "peer"
    send_counts = (at_bounds[rank, 1:] - at_bounds[rank, :-1]).astype(np.int64)
    recv_counts = (at_bounds[:, rank + 1] - at_bounds[:, rank]).astype(np.int64)
    n_recv = int(recv_counts.sum())
    words = 1 if eb == 4 else eb // 8
    tstr, tdt = ("<i4", torch.int32) if eb == 4 else ("<i8", torch.int64)
    send = _wrap(elems_ptr, max(n_elems, 1) * words, tstr, dev) if elems_ptr else torch.empty(1, dtype=tdt, device=dev)
    need = max(n_recv, 1) * eb
    buf = getattr(kc, "_recv_buf", None)
    if buf is None or buf.numel() * 8 < need:
        kc._recv_buf = None
        del buf
        torch.cuda.empty_cache()
        buf = torch.empty(int(need * 1.02) // 8 + 1024, dtype=torch.int64, device=dev)
        kc._recv_buf = buf
    recv = buf.view(tdt)
    dist.all_to_all_single(recv[: n_recv * words], send[: n_elems * words],
                           output_split_sizes=[int(c) * words for c in recv_counts],
                           input_split_sizes=[int(c) * words for c in send_counts], group=group)
    torch.cuda.current_stream().synchronize()
    del send
    seg_off = np.concatenate([[0], np.cumsum(recv_counts)[:-1]]).astype(np.uint64)
    kc.count_pieces(recv.data_ptr(), world, sizes_u32.data_ptr(), seg_off, lo, hi)
    return lo, hi, "nccl"


def _sharded_count_rounds(kc, rank, world, P, dev, group, timings, on_round, want_peer):
    """The partition-first pipeline in k-mer-space rounds (see the module docstring)."""
    tot0, cap = kc.level0_totals()
    t = torch.from_numpy(np.concatenate([tot0.astype(np.int64), [-(cap if cap else 2 ** 62)]])).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)   # per bucket: the largest share; smallest capacity (as -cap)
    t = t.cpu().numpy()
    cap_all = int(-t[-1])
    forced = int(os.environ.get("APGK_SHARD_ROUND_KEYS", "0"))
    # a round holds this rank's partition AND the shard it receives (about the same size): half the budget each
    capacity = forced if forced > 0 else max(1, cap_all // 2)
    rounds = plan_rounds(t[:-1], capacity)
    spec = {}
    ni = nd = 0
    path = "?"
    ranges = []
    for i, (lo0, hi0) in enumerate(rounds):
        kc.partition_range(P, lo0, hi0)
        blo, bhi, path = _exchange_and_count(kc, rank, world, P, dev, group, want_peer)
        ranges.append((blo, bhi))
        a, b = kc.totals()
        ni += a
        nd += b
        f, m = kc.spectrum_sparse()
        for x, y in zip(f.tolist(), m.tolist()):
            spec[x] = spec.get(x, 0) + y
        if on_round is not None:
            on_round(kc, i, len(rounds))
        dist.barrier(group=group)   # the peers are done reading this round's partition buffers
    out, ni_g, nd_g = merge_sparse_spectra(spec, ni, nd, world, dev, group)
    if timings is not None:
        timings["path"] = "partition-first/rounds/" + path
        timings["n_rounds"] = len(rounds)
        timings["prefix_bits"] = int(P)
        timings["level0_rounds"] = [tuple(r) for r in rounds]
        timings["bucket_ranges"] = ranges
    return out, ni_g, nd_g


def merge_sparse_spectra(spec, ni, nd, world, dev, group=None):
    """Sum per-rank sparse spectra {frequency: n_kmers} (disjoint k-mer sets, so plain sums) and the totals over
    the ranks: frequencies below 65536 and the totals in one all-reduce, the (rare) larger ones gathered.
    -> (dense uint64 spectrum up to the largest frequency, n_instances, n_distinct).  Works on gloo and nccl."""
    DENSE = 65536
    dense = np.zeros(DENSE + 2, dtype=np.int64)
    far = {}
    for x, y in spec.items():
        if x < DENSE:
            dense[x] = y
        else:
            far[int(x)] = int(y)
    dense[DENSE], dense[DENSE + 1] = ni, nd
    td = torch.from_numpy(dense).to(dev)
    dist.all_reduce(td, op=dist.ReduceOp.SUM, group=group)
    dense = td.cpu().numpy()
    far_all = [None] * world
    dist.all_gather_object(far_all, far, group=group)
    far = {}
    for d_ in far_all:
        for x, y in d_.items():
            far[x] = far.get(x, 0) + y
    nz = np.nonzero(dense[:DENSE])[0]
    top = max([int(nz.max()) if len(nz) else 0] + list(far.keys()))
    out = np.zeros(top + 1, dtype=np.uint64)
    out[: min(top + 1, DENSE)] = dense[: min(top + 1, DENSE)].astype(np.uint64)
    for x, y in far.items():
        out[x] = y
    return out, int(dense[DENSE]), int(dense[DENSE + 1])


def _open_peers(kc, rank, world, handles):
    """Device pointers of every rank's partition buffer (None for this rank's own), mapping the peers'
    IPC handles once and again only when a peer reallocated.  None if a handle cannot be mapped."""
    from .kmers import ApgkError

    cache = getattr(kc, "_peer_map", None)
    if cache is None:
        cache = kc._peer_map = {}
    ptrs = []
    for s in range(world):
        if s == rank:
            ptrs.append(None)
            continue
        key = handles[s].tobytes()
        ent = cache.get(s)
        if ent is None or ent[0] != key:
            try:
                if ent is not None:
                    kc.peer_close(ent[1])
                    del cache[s]
                ptr = kc.peer_open(np.frombuffer(key, dtype=np.uint8))
            except ApgkError:
                return None
            cache[s] = (key, ptr)
        ptrs.append(cache[s][1])
    return ptrs


def _reduce_results(kc, dev, group):
    """sum the dense spectra (and the totals, in the same all-reduce) on the device; the library then
    reloads its host copy of the spectrum"""
    ptr, n = kc.spectrum_device()
    spec = _wrap(ptr, n, "<i8", dev)
    ni, nd = kc.totals()
    both = torch.cat([spec, torch.tensor([ni, nd], dtype=torch.int64, device=dev)])
    dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
    spec.copy_(both[:n])
    tot = both[n:].cpu()
    kc.spectrum_reload()
    return kc.spectrum(), int(tot[0]), int(tot[1])


def exchange_plan(send_counts, world, group=None, device=None):
    """All ranks learn how many k-mers they receive from every peer.
    send_counts: uint64[world] (k-mers this rank sends to each owner) -> recv_counts int64[world]."""
    t_send = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=device)
    t_recv = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(t_recv, t_send, group=group)
    return t_recv.cpu().numpy()


def sharded_count_hash(kc, rank, world, group=None, timings=None):
    """Hash-owner form of the sharded pipeline (see the module docstring); same contract as sharded_count."""
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    W = kc.W
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    send_counts = kc.owner_plan(world)
    n_send = int(send_counts.sum())
    # send buffer = the library's own level-0 key buffer (it is rewritten by the count below anyway);
    # receive buffer = one torch tensor kept on the counter and reused from step to step
    send = _wrap_u64(kc.key_buffer(n_send), max(n_send, 1) * W, dev)
    kc.owner_scatter(send.data_ptr())  # synchronous: the library's stream is drained on return
    recv_counts = exchange_plan(send_counts, world, group, dev)
    n_recv = int(recv_counts.sum())
    recv = getattr(kc, "_recv_buf", None)
    if recv is None or recv.numel() < n_recv * W:
        kc._recv_buf = None
        del recv
        torch.cuda.empty_cache()
        recv = torch.empty(int(max(n_recv, 1) * W * 1.02) + 1024, dtype=torch.int64, device=dev)
        kc._recv_buf = recv
    e1.record()
    dist.all_to_all_single(recv[: n_recv * W], send[: n_send * W],
                           output_split_sizes=[int(c) * W for c in recv_counts],
                           input_split_sizes=[int(c) * W for c in send_counts], group=group)
    e2.record()
    torch.cuda.current_stream().synchronize()
    del send
    kc.finish_keys_device(recv.data_ptr(), n_recv)
    # sum the dense spectra in place on the device, then reload on the host side of the library
    ptr, n = kc.spectrum_device()
    spec = _wrap_u64(ptr, n, dev)
    dist.all_reduce(spec, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream().synchronize()
    kc.spectrum_reload()
    ni, nd = kc.totals()
    tot = torch.tensor([ni, nd], dtype=torch.int64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    if timings is not None:
        timings.setdefault("path", "hash")
        timings["owner_ms"] = e0.elapsed_time(e1)
        timings["all_to_all_ms"] = e1.elapsed_time(e2)
        timings["sent_kmers"] = n_send
        timings["recv_kmers"] = n_recv
    return kc.spectrum(), int(tot[0].item()), int(tot[1].item())


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can view library-owned device memory."""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _wrap(ptr, n, typestr, dev):
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device=dev)


def _wrap_u64(ptr, n, dev):
    return _wrap(ptr, n, "<i8", dev)


# ---------------------------------------------------------------------------
# Host-side logic of the shuffle, separated so it can be tested on CPU with gloo:
# the same ownership rule and the same exchange, with numpy doing what the kernels do.
# ---------------------------------------------------------------------------
def host_shuffle(kmers, K, rank, world, group=None):
    """kmers: uint64[n, W] canonical k-mer instances held by this rank (host).  Routes every
    instance to its owner with all_to_all (gloo or nccl) and returns the instances this rank owns."""
    from .kmers import owner_of, words_per_kmer

    W = words_per_kmer(K)
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, W)
    own = owner_of(K, kmers, world) if len(kmers) else np.zeros(0, dtype=np.uint32)
    order = np.argsort(own, kind="stable")
    send_counts = np.bincount(own, minlength=world).astype(np.int64)
    send = torch.from_numpy(kmers[order].astype(np.int64).reshape(-1))
    t_recv_counts = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(t_recv_counts, torch.from_numpy(send_counts), group=group)
    recv_counts = t_recv_counts.numpy()
    recv = torch.empty(int(recv_counts.sum()) * W, dtype=torch.int64)
    dist.all_to_all_single(recv, send, output_split_sizes=[int(c) * W for c in recv_counts],
                           input_split_sizes=[int(c) * W for c in send_counts], group=group)
    return recv.numpy().astype(np.uint64).reshape(-1, W)


def host_partition_shuffle(kmers, K, prefix_bits, rank, world, group=None):
    """Host mirror of the partition-first exchange (gloo or nccl): kmers uint64[n, W] canonical k-mer
    instances held by this rank.  Buckets = leading prefix_bits of the 2K-bit k-mer; bucket ranges are
    cut by `balanced_splitters` from the all-gathered histogram; returns (the instances this rank owns,
    (bucket_lo, bucket_hi))."""
    from .kmers import words_per_kmer

    W = words_per_kmer(K)
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, W)
    top = 2 * K - 64 * (W - 1)  # bits in word 0
    if prefix_bits > 2 * K:
        raise ValueError("prefix longer than the k-mer")
    if prefix_bits <= top:
        bucket = (kmers[:, 0] >> np.uint64(top - prefix_bits)).astype(np.int64) if len(kmers) else np.zeros(0, np.int64)
    else:  # the prefix reaches into word 1
        r = prefix_bits - top
        bucket = ((kmers[:, 0] << np.uint64(r)) | (kmers[:, 1] >> np.uint64(64 - r))).astype(np.int64)
    nb = 1 << prefix_bits
    sizes = torch.from_numpy(np.bincount(bucket, minlength=nb).astype(np.int64))
    all_sizes = [torch.empty(nb, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes)
    bounds = balanced_splitters(all_sizes.sum(0), world)
    order = np.argsort(bucket, kind="stable")
    cum = np.concatenate([[0], np.cumsum(all_sizes.numpy(), axis=1)[rank]])
    send_counts = np.array([cum[bounds[r + 1]] - cum[bounds[r]] for r in range(world)], dtype=np.int64)
    send = torch.from_numpy(kmers[order].astype(np.int64).reshape(-1))
    t_recv_counts = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(t_recv_counts, torch.from_numpy(send_counts), group=group)
    recv_counts = t_recv_counts.numpy()
    recv = torch.empty(int(recv_counts.sum()) * W, dtype=torch.int64)
    dist.all_to_all_single(recv, send, output_split_sizes=[int(c) * W for c in recv_counts],
                           input_split_sizes=[int(c) * W for c in send_counts], group=group)
    return recv.numpy().astype(np.uint64).reshape(-1, W), (bounds[rank], bounds[rank + 1])
