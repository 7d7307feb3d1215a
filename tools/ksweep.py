"""BASELINE.json configs[4]: K sweep 20/25/48/64/96 (1-, 2- and 3-word k-mers) at 1/2/4/8 GPUs -- 24 M x 250 bp reads of an
N x 100 Mb genome per GPU (weak scaling), resident reads, one JSON line per K.

   python tools/ksweep.py [reads_per_gpu]                                          (one GPU)
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 tools/ksweep.py

At N > 1 the count is apgk_group_count (dist.sharded_count); exact invariants are checked for every K:
sum f * spectrum[f] == instances == reads * (L - K + 1), sum spectrum == distinct.  `lsd_model_ratio` is the contract figure of
SURVEY.md section 8(d) (B_alg(K) bytes per instance against the measured HBM peak of all GPUs), not an achieved bandwidth."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from allpathslg_b200 import KmerCounter, synth_params

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24_000_000
G, L = 100_000_000 * world, 250
dist = None
if world > 1:
    import torch
    import torch.distributed as dist

    from allpathslg_b200.dist import sharded_count

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    HBM = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    HBM = 6451.5
for K in (20, 25, 48, 64, 96):
    kc = KmerCounter(K, device=local)
    kc.synth_reads(synth_params(G, L), rank * n, n)

    def step():
        if world == 1:
            kc.finish()
            return kc.spectrum(), kc.totals()[0], kc.totals()[1], {}
        tm = {}
        s, a, b = sharded_count(kc, rank, world, timings=tm)
        return s, a, b, tm

    step()  # warm-up (allocations, mappings)
    reps, dts = 2, []
    for _ in range(reps):
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
        t0 = time.perf_counter()
        spec, ni, nd, tm = step()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
        dts.append(time.perf_counter() - t0)
    dt = min(dts)
    if dist is not None:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    ok = int((spec * np.arange(len(spec), dtype=np.uint64)).sum()) == ni == n * world * (L - K + 1) and int(spec.sum()) == nd
    S, P = 8 * ((2 * K + 63) // 64), (2 * K + 7) // 8
    b_alg = S * (2 * P + 3)
    if rank == 0:
        st = {k: round(v, 2) for k, v in kc.stage_ms().items() if v > 0}
        print(json.dumps({"K": K, "words": kc.W, "n_gpus": world, "reads_per_gpu": n, "read_len": L, "instances": ni, "distinct": nd,
                          "Gkmers_per_s": round(ni / dt / 1e9, 2), "ms": round(dt * 1e3, 1), "invariants_ok": bool(ok),
                          "lsd_model_ratio": round(ni / dt * b_alg / (world * HBM * 1e9), 3),
                          "geometry": kc.geometry(), "rounds": tm.get("n_rounds", kc.geometry()["n_rounds"]), "stage_ms": st}), flush=True)
    if dist is not None:
        kc._group.close()
        dist.barrier()
    kc.close()
if dist is not None:
    dist.destroy_process_group()
