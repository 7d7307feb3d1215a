"""torchrun script: sharded multi-GPU count vs the oracle on the union of all ranks' reads.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from allpathslg_b200 import KmerCounter, synth_params, owner_of
from allpathslg_b200.dist import sharded_count
from oracle import oracle_a as A

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (K, G, L, n_per) in [(25, 2_000_000, 100, 400_000), (64, 500_000, 250, 40_000), (96, 500_000, 150, 60_000), (20, 300_000, 100, 100_000)]:
    kc = KmerCounter(K, device=local)
    kc.synth_reads(synth_params(G, L), rank * n_per, n_per)
    tm = {}
    spec, ni, nd = sharded_count(kc, rank, world, timings=tm)
    # this rank's shard table must be exactly the oracle's k-mers it owns
    packed, off = A.synth_reads(A.synth_params(G, L), 0, n_per * world)
    ek, ec, en = A.count(packed, off, K)
    es = A.spectrum(ec)
    gk, gc = kc.counts()
    if tm["path"] == "hash":
        sel = owner_of(K, ek, world) == rank
    else:  # partition-first: this rank owns a range of the P-bit prefix buckets
        P = tm["prefix_bits"]  # the exchange's bucket space (the shard's own table is indexed finer)
        W = ek.shape[1]
        top = 2 * K - 64 * (W - 1)
        assert P <= top
        bucket = (ek[:, 0] >> np.uint64(top - P)).astype(np.int64)
        lo, hi = tm["bucket_range"]
        sel = (bucket >= lo) & (bucket < hi)
    good = (ni == en and nd == len(ek) and len(spec) == len(es) and (spec == es).all()
            and len(gk) == int(sel.sum()) and (gk == ek[sel]).all() and (gc.astype(np.uint64) == ec[sel]).all())
    t = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("K=%d world=%d instances=%d distinct=%d all-ranks-ok=%s  (%s; rank0 sent %d recv %d, a2a %.2f ms)" %
              (K, world, ni, nd, bool(t.item()), tm["path"], tm.get("sent_elems", tm.get("sent_kmers", 0)),
               tm.get("recv_elems", tm.get("recv_kmers", 0)), tm["all_to_all_ms"]), flush=True)
    ok = ok and bool(t.item())
    kc.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
