"""BASELINE.json configs[3]: synthetic human-size genome (3 Gb) at 45x, K=25 spectrum sharded across the GPUs of one box
with streamed batches and k-mer-space rounds.  torchrun script:

   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
       tools/human_scale.py [--genome 3000000000] [--coverage 45] [--read-len 100] [--batch-reads 20000000] [--K 25]

Every rank generates its share of the reads on the device in batches (the stand-in for batches streamed from pinned
host memory: `--host-batches` really stages each batch through a pinned host buffer and apgk_add_reads_uniform),
then dist.sharded_count runs the partition-first pipeline -- in k-mer-space rounds when the rank's k-mers do not fit
the device at once.  The result is checked with the size-independent invariants (sum f * spectrum[f] == instances,
per-rank totals add up) and one JSON line is printed by rank 0.  Not run at full size in round 1 (no 8-GPU budget);
`--genome 30000000 --batch-reads 2000000` is the smoke configuration."""
import argparse, json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from allpathslg_b200 import KmerCounter, synth_params
from allpathslg_b200.dist import sharded_count

ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=3_000_000_000)
ap.add_argument("--coverage", type=float, default=45.0)
ap.add_argument("--read-len", type=int, default=100)
ap.add_argument("--batch-reads", type=int, default=20_000_000)
ap.add_argument("--K", type=int, default=25)
ap.add_argument("--host-batches", action="store_true")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L, K = args.read_len, args.K
n_total = int(args.genome * args.coverage / L)
n_mine = n_total // world + (1 if rank < n_total % world else 0)
first = rank * (n_total // world) + min(rank, n_total % world)
sp = synth_params(args.genome, L)
kc = KmerCounter(K, device=local, want_counts=True, reserve_bases=n_mine * L)
t0 = time.perf_counter()
stage = KmerCounter(K, device=local, want_counts=False) if args.host_batches else None
pinned = torch.empty(((args.batch_reads * L + 31) // 32) * 8, dtype=torch.uint8).pin_memory() if args.host_batches else None
done = 0
while done < n_mine:
    n = min(args.batch_reads, n_mine - done)
    if args.host_batches:       # generate on a scratch context, bring the batch to pinned host memory, stream it in
        stage.reset(); stage.synth_reads(sp, first + done, n); stage.export_reads(pinned.data_ptr())
        kc.add_reads_uniform(pinned.data_ptr(), n, L)
    else:
        kc.synth_reads(sp, first + done, n)
    done += n
torch.cuda.synchronize(); dist.barrier()
t_ingest = time.perf_counter() - t0
tm = {}
t0 = time.perf_counter()
spec, ni, nd = sharded_count(kc, rank, world, timings=tm)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
expect = n_total * (L - K + 1)
ok = ni == expect and int((spec * np.arange(len(spec), dtype=np.uint64)).sum()) == ni and int(spec.sum()) == nd
if rank == 0:
    print(json.dumps({"workload": "synthetic %.2f Gb genome at %gx, %d x %d bp reads, K=%d, %d GPUs" % (args.genome / 1e9, args.coverage, n_total, L, K, world),
                      "n_instances": ni, "n_distinct": nd, "invariants_ok": bool(ok), "count_s": round(dt, 3), "ingest_s": round(t_ingest, 3),
                      "Gkmers_per_s": round(ni / dt / 1e9, 2), "path": tm.get("path"), "n_rounds": tm.get("n_rounds", 1),
                      "prefix_bits": tm.get("prefix_bits")}), flush=True)
kc.close()
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
