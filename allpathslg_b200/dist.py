"""Multi-GPU k-mer spectrum: one process per GPU, hash-sharded by canonical k-mer.

Each rank extracts the canonical k-mers of ITS reads and groups them by owner
rank (owner = hash(k-mer) % world, `apgk_owner_plan` / `apgk_owner_scatter`),
the groups are exchanged with one NCCL all-to-all over NVLink
(`torch.distributed.all_to_all_single`), each rank sorts and counts the shard it
owns (`apgk_finish_keys_device`), and the per-rank spectra -- disjoint sets of
k-mers, so plain integer sums -- are all-reduced.  Identical k-mers always land
on the same rank, so counts are final without a merge.

torch is used for device buffers, streams and the collective only.
"""
import numpy as np
import torch
import torch.distributed as dist


def exchange_plan(send_counts, world, group=None, device=None):
    """All ranks learn how many k-mers they receive from every peer.
    send_counts: uint64[world] (k-mers this rank sends to each owner) -> recv_counts int64[world]."""
    t_send = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=device)
    t_recv = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(t_recv, t_send, group=group)
    return t_recv.cpu().numpy()


def sharded_count(kc, rank, world, group=None, timings=None):
    """Run the sharded pipeline on this rank's KmerCounter `kc` (reads already in its store).

    Returns (spectrum uint64 array summed over all ranks, n_instances_global, n_distinct_global).
    The rank's own shard table stays queryable in `kc` (counts of the k-mers it owns)."""
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
        timings["owner_ms"] = e0.elapsed_time(e1)
        timings["all_to_all_ms"] = e1.elapsed_time(e2)
        timings["sent_kmers"] = n_send
        timings["recv_kmers"] = n_recv
    return kc.spectrum(), int(tot[0].item()), int(tot[1].item())


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can view library-owned device memory."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def _wrap_u64(ptr, n, dev):
    return torch.as_tensor(_CudaArray(ptr, n), device=dev)


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
