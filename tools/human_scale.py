"""BASELINE.json configs[3]: synthetic human-size genome (3 Gb) at 45x, K=25 spectrum + counts sharded across the GPUs
of one box with streamed batches and k-mer-space rounds -- the north_star's target run, with its parity check.

   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
       tools/human_scale.py [--genome 3000000000] [--coverage 45] [--read-len 100] [--batch-reads 20000000] [--K 25] \
                            [--steps 2] [--out profiles/r02_human_scale.json]

Every rank
  1. generates its share of the reads batch by batch (device generator, same function as the oracle's) into ONE pinned
     host buffer -- from here on the reads are host data, as they would be for the reference;
  2. streams the batches from pinned host memory into its context (apgk_add_reads_uniform);
  3. takes part in apgk_group_count (dist.sharded_count): levels 0 + 1, exchange over NVLink peer memory, counting,
     spectrum all-reduce, in k-mer-space rounds (outer / inner) sized from the device memory left;
  4. checks the size-independent invariants (sum f * spectrum[f] == instances == reads * (L - K + 1), sum spectrum ==
     distinct == sum of the shard tables' sizes, every shard table strictly ascending);
  5. SAMPLED-PARTITION PARITY (SURVEY.md section 8c "human-scale check" ii): the CPU oracle scans ALL of the rank's host
     reads and keeps the k-mer instances of 4 of 256 leading-bit partitions (oracle_sample_prefix); the instances of all
     ranks are brought together per partition (all-gather), counted exactly (oracle_count_keys) and compared,
     record by record, with the records the GPU shards hold for the same key ranges (apgk_prefix_range).
One JSON line is printed by rank 0 (and written to --out).  The oracle is the spec-derived restatement: parity unpinned.
`--genome 30000000 --batch-reads 2000000` is the smoke configuration."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

from allpathslg_b200 import KmerCounter, synth_params
from allpathslg_b200.dist import sharded_count
from oracle import oracle_a as A

ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=3_000_000_000)
ap.add_argument("--coverage", type=float, default=45.0)
ap.add_argument("--read-len", type=int, default=100)
ap.add_argument("--batch-reads", type=int, default=20_000_000)
ap.add_argument("--K", type=int, default=25)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--pbits", type=int, default=8)
ap.add_argument("--parts", type=str, default="5,77,130,201")
ap.add_argument("--no-parity", action="store_true")
ap.add_argument("--out", type=str, default="")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L, K = args.read_len, args.K
parts = [int(x) for x in args.parts.split(",")]
n_total = int(args.genome * args.coverage / L)
n_mine = n_total // world + (1 if rank < n_total % world else 0)
first = rank * (n_total // world) + min(rank, n_total % world)
sp = synth_params(args.genome, L)
assert (args.batch_reads * L) % 32 == 0, "batches must end on an 8-byte boundary of the packed stream"
assert K <= 32, "the sampled-partition check of this tool handles one-word k-mers"


def log(msg):
    if rank == 0:
        print("[human_scale %7.1fs] %s" % (time.perf_counter() - T0, msg), file=sys.stderr, flush=True)


T0 = time.perf_counter()
# ---- 1. the reads, as host data
host = torch.empty(((n_mine * L + 31) // 32) * 8, dtype=torch.uint8).pin_memory()
stage = KmerCounter(K, device=local, want_counts=False, reserve_bases=min(args.batch_reads, n_mine) * L)
done = 0
while done < n_mine:
    n = min(args.batch_reads, n_mine - done)
    stage.reset()
    stage.synth_reads(sp, first + done, n)
    stage.export_reads(host.data_ptr() + done * L // 4)
    done += n
stage.close()
del stage
torch.cuda.synchronize()
dist.barrier()
t_gen = time.perf_counter() - T0
log("reads generated: %d per rank, %.2f GB packed per rank" % (n_mine, host.numel() / 1e9))

# ---- 2. streamed ingest from pinned host memory
kc = KmerCounter(K, device=local, want_counts=True, reserve_bases=n_mine * L)
t0 = time.perf_counter()
done = 0
while done < n_mine:
    n = min(args.batch_reads, n_mine - done)
    kc.add_reads_uniform(host.data_ptr() + done * L // 4, n, L)
    done += n
torch.cuda.synchronize()
dist.barrier()
t_ingest = time.perf_counter() - t0
log("ingest done in %.2f s" % t_ingest)
# room for the shard's table: ~0.15 distinct k-mers per instance at 45x with 0.5 % errors, plus margin
kc.reserve_table(int(n_mine * (L - K + 1) * 0.21) + (1 << 20))

# ---- 3. the count (first step sizes and maps the buffers; the following ones are timed)
times, tm = [], {}
spec = ni = nd = None
for step in range(1 + args.steps):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    tm = {}
    spec, ni, nd = sharded_count(kc, rank, world, timings=tm)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
    log("step %d: %.3f s (%s, %d outer / %d inner rounds, P=%d+%d)" % (step, times[-1], tm["path"], tm["n_outer_rounds"], tm["n_rounds"],
                                                                     tm["prefix_bits"], tm["split_bits"]))
best = min(times[1:]) if len(times) > 1 else times[0]

# ---- 4. invariants
expect = n_total * (L - K + 1)
n_shard, d_shard = kc.totals()
tot = torch.tensor([n_shard, d_shard], dtype=torch.int64, device="cuda")
dist.all_reduce(tot)
inv = {
    "instances_eq_reads_x_windows": ni == expect,
    "sum_f_spectrum_eq_instances": int((spec * np.arange(len(spec), dtype=np.uint64)).sum()) == ni,
    "sum_spectrum_eq_distinct": int(spec.sum()) == nd,
    "shard_totals_add_up": int(tot[0]) == ni and int(tot[1]) == nd,
}

# ---- 5. sampled-partition parity
parity = None
if not args.no_parity:
    def gather_all(arr):
        """every rank's 1-D array (uint64 / uint32) on every GPU; -> list of device tensors (views as int64 / int32)"""
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64 if arr.dtype == np.uint64 else np.int32)).cuda()
        n = torch.tensor([t.numel()], dtype=torch.int64, device="cuda")
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        mx = max(1, max(int(x) for x in ns))
        pad = torch.zeros(mx, dtype=t.dtype, device="cuda")
        pad[: t.numel()] = t
        out = torch.empty(world * mx, dtype=t.dtype, device="cuda")
        dist.all_gather_into_tensor(out, pad)
        return [out[r * mx: r * mx + int(ns[r])] for r in range(world)]

    kc.release_temp()   # the count's scratch buffers are not needed any more: room for the gathers below
    ncpu = os.cpu_count() or 1
    threads = max(1, ncpu // world)
    t0 = time.perf_counter()
    keys, n_win = A.sample_prefix(host.data_ptr(), None, K, args.pbits, parts, n_reads=n_mine, read_len=L, n_threads=threads)
    t_scan = time.perf_counter() - t0
    top = 2 * K
    pre = (keys[:, 0] >> np.uint64(top - args.pbits)).astype(np.int64)
    mine_ok, n_rec, n_inst, t_cnt = 1, 0, 0, 0.0
    for j, part in enumerate(parts):
        # the oracle's instances and the device's records of the partition, from every rank, to the rank that checks it
        f0, n = kc.prefix_range(args.pbits, part)
        gk, gc = kc.counts(f0, n)
        cpu_all = gather_all(keys[pre == part, 0])
        gk_all = gather_all(gk[:, 0] if len(gk) else np.zeros(0, np.uint64))
        gc_all = gather_all(gc)
        if j % world == rank:
            allk = torch.cat(cpu_all).cpu().numpy().view(np.uint64)
            gk = torch.cat(gk_all).cpu().numpy().view(np.uint64)
            gc = torch.cat(gc_all).cpu().numpy().view(np.uint32)
        del cpu_all, gk_all, gc_all
        torch.cuda.empty_cache()
        if j % world != rank:
            continue
        t1 = time.perf_counter()
        ek, ec = A.count_keys(allk.reshape(-1, 1), K, n_threads=ncpu)
        t_cnt += time.perf_counter() - t1
        order = np.argsort(gk, kind="stable")   # a rank's slice ascends; the ranks' slices interleave round by round
        good = len(gk) == len(ek) and bool((gk[order] == ek[:, 0]).all()) and bool((gc[order].astype(np.uint64) == ec).all())
        mine_ok = mine_ok and int(good)
        n_rec += len(ek)
        n_inst += len(allk)
        del allk, gk, gc, ek, ec, order
    del keys, pre
    res = torch.tensor([mine_ok, n_rec, n_inst, int(n_win)], dtype=torch.int64, device="cuda")
    mn = res.clone()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(res)
    tt = torch.tensor([t_scan, t_cnt], device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    parity = {"what": "CPU oracle over ALL reads restricted to %d of %d leading-bit partitions vs the device shards' records of the "
                      "same key ranges, record by record" % (len(parts), 1 << args.pbits),
              "ok": bool(int(mn[0])), "records_compared": int(res[1]), "instances_in_sample": int(res[2]),
              "windows_scanned": int(res[3]), "windows_scanned_eq_instances": int(res[3]) == ni,
              "cpu_scan_s": round(float(tt[0]), 2), "cpu_count_s": round(float(tt[1]), 2), "cpu_threads_per_rank": threads,
              "oracle": "spec-derived restatement (oracle/kmer_oracle.c); the reference source was not available: parity unpinned"}
    dist.barrier()

ok = all(inv.values()) and (parity is None or (parity["ok"] and parity["windows_scanned_eq_instances"]))
if rank == 0:
    hbm = 6451.5
    try:
        hbm = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        pass
    gk_s = ni / best / 1e9
    line = {"workload": "synthetic %.2f Gb genome at %gx, %d x %d bp reads, K=%d, %d GPUs, batches of %d reads streamed from pinned host memory"
                        % (args.genome / 1e9, args.coverage, n_total, L, K, world, args.batch_reads),
            "n_instances": int(ni), "n_distinct": int(nd), "invariants": inv, "invariants_ok": bool(all(inv.values())),
            "parity_sample": parity, "ok": bool(ok),
            "count_s": round(best, 4), "count_s_all_steps": [round(x, 4) for x in times], "Gkmers_per_s": round(gk_s, 2),
            "lsd_model_ratio_of_aggregate_hbm": round(gk_s * 136.0 / (world * hbm), 4),
            "target_Gkmers_per_s": round(0.5 * world * hbm / 136.0, 1),
            "ingest_s": round(t_ingest, 3), "generate_s": round(t_gen, 2),
            "path": tm.get("path"), "n_rounds": tm.get("n_rounds"), "n_outer_rounds": tm.get("n_outer_rounds"),
            "prefix_bits": tm.get("prefix_bits"), "split_bits": tm.get("split_bits"),
            "rank0": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()},
            "spectrum_head": [int(x) for x in spec[:64]]}
    print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write(json.dumps(line, indent=1) + "\n")
kc._group.close()
dist.barrier()   # every rank has unmapped the peers' buffers before anyone frees its own
kc.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
