"""BASELINE.json configs[4]: K sweep 20/25/48/64/96 on the 100 Mb genome, 24 M x 250 bp reads (1 GPU).
Prints one JSON line per K: throughput with resident reads, stage times, geometry, exact invariants."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
from allpathslg_b200 import KmerCounter, synth_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24_000_000
G, L = 100_000_000, 250
for K in (20, 25, 48, 64, 96):
    kc = KmerCounter(K)
    kc.synth_reads(synth_params(G, L), 0, n)
    kc.finish()  # warm-up (allocations)
    t0 = time.perf_counter(); reps = 2
    for _ in range(reps):
        kc.finish()
    dt = (time.perf_counter() - t0) / reps
    ni, nd = kc.totals()
    spec = kc.spectrum()
    ok = int((spec * np.arange(len(spec), dtype=np.uint64)).sum()) == ni == n * (L - K + 1) and int(spec.sum()) == nd
    st = {k: round(v, 2) for k, v in kc.stage_ms().items() if v > 0}
    print(json.dumps({"K": K, "words": kc.W, "reads": n, "read_len": L, "instances": ni, "distinct": nd,
                      "Gkmers_per_s": round(ni / dt / 1e9, 2), "ms": round(dt * 1e3, 1), "invariants_ok": bool(ok),
                      "geometry": kc.geometry(), "stage_ms": st}), flush=True)
    kc.close()
