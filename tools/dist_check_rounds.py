"""torchrun script: the sharded count in K-MER-SPACE ROUNDS (a small APGK_ROUND_KEYS makes apgk_partition refuse a
single round) vs the oracle on the union of all ranks' reads.
   APGK_ROUND_KEYS=... python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
       --master-port 29513 tools/dist_check_rounds.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from allpathslg_b200 import KmerCounter, synth_params
from allpathslg_b200.dist import sharded_count
from oracle import oracle_a as A

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (K, G, L, n_per) in [(25, 2_000_000, 100, 400_000), (48, 500_000, 150, 60_000)]:
    os.environ["APGK_ROUND_KEYS"] = str(n_per * (L - K + 1) // 3)   # a third of a rank's instances per round
    kc = KmerCounter(K, device=local)
    kc.synth_reads(synth_params(G, L), rank * n_per, n_per)
    tm, tabs = {}, []
    spec, ni, nd = sharded_count(kc, rank, world, timings=tm, on_round=lambda c, i, n: tabs.append(c.counts()))
    packed, off = A.synth_reads(A.synth_params(G, L), 0, n_per * world)
    ek, ec, en = A.count(packed, off, K)
    es = A.spectrum(ec)
    good = tm["path"].startswith("partition-first/rounds") and tm["n_rounds"] >= 2
    good = good and ni == en and nd == len(ek) and len(spec) == len(es) and bool((spec == es).all())
    P = tm["prefix_bits"]; D0 = (P + 1) // 2; D1 = P - D0
    W = ek.shape[1]; top = 2 * K - 64 * (W - 1)
    bucket = (ek[:, 0] >> np.uint64(top - P)).astype(np.int64)
    for (lo0, hi0), (blo, bhi), (gk, gc) in zip(tm["level0_rounds"], tm["bucket_ranges"], tabs):
        a, b = max(blo, lo0 << D1), min(bhi, hi0 << D1)
        sel = (bucket >= a) & (bucket < b)
        good = good and len(gk) == int(sel.sum()) and bool((gk == ek[sel]).all()) and bool((gc.astype(np.uint64) == ec[sel]).all())
    t = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("K=%d world=%d rounds=%d instances=%d distinct=%d all-ranks-ok=%s (%s)" % (K, world, tm["n_rounds"], ni, nd, bool(t.item()), tm["path"]), flush=True)
    ok = ok and bool(t.item())
    kc.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST ROUNDS CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
