"""Cost of the sharded form's k-mer-space rounds at the bench workload on ONE rank (torchrun, world 1):
the same reads counted in one round and in forced rounds (APGK_ROUND_KEYS)."""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from allpathslg_b200 import KmerCounter, synth_params
from allpathslg_b200.dist import sharded_count
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
kc = KmerCounter(25, device=0, want_counts=True)
kc.synth_reads(synth_params(100_000_000, 100), 0, 60_000_000)
for cap in (0, 2_400_000_000, 1_200_000_000):
    if cap:
        os.environ["APGK_ROUND_KEYS"] = str(cap)
    for rep in range(3):
        tm = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        spec, ni, nd = sharded_count(kc, 0, 1, timings=tm)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("round capacity %s: %s, %d rounds, %.1f ms, %.1f Gk-mers/s" % (cap or "auto", tm["path"], tm.get("n_rounds", 1), dt * 1e3, ni / dt / 1e9), flush=True)
dist.destroy_process_group()
