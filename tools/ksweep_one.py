"""One K of the K sweep (argv: K [reads]); prints throughput, stage times, geometry.  Env knobs apply (APGK_PREFIX_BITS ...)."""
import json, sys, time, os
sys.path.insert(0, ".")
import numpy as np
from allpathslg_b200 import KmerCounter, synth_params
K = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 24_000_000
G, L = 100_000_000, 250
kc = KmerCounter(K)
kc.synth_reads(synth_params(G, L), 0, n)
kc.finish()
t0 = time.perf_counter(); reps = 2
for _ in range(reps):
    kc.finish()
dt = (time.perf_counter() - t0) / reps
ni, nd = kc.totals()
spec = kc.spectrum()
ok = int((spec * np.arange(len(spec), dtype=np.uint64)).sum()) == ni == n * (L - K + 1) and int(spec.sum()) == nd
print(K, os.environ.get("APGK_PREFIX_BITS"), round(ni / dt / 1e9, 2), "Gk/s", round(dt * 1e3, 1), "ms", ok, kc.geometry(),
      {k: round(v) for k, v in kc.stage_ms().items() if v > 2}, flush=True)
