"""host-side stage trace of one finish (APGK_TRACE=1): where does the wall clock go for a given K?"""
import sys, time
sys.path.insert(0, ".")
from allpathslg_b200 import KmerCounter, synth_params
K = int(sys.argv[1]); n = int(sys.argv[2]); L = int(sys.argv[3]) if len(sys.argv) > 3 else 250
kc = KmerCounter(K)
kc.synth_reads(synth_params(100_000_000, L), 0, n)
kc.finish()
kc.debug_counters()
print("---- second finish", file=sys.stderr, flush=True)
t0 = time.perf_counter(); kc.finish(); dt = time.perf_counter() - t0
print("K", K, "ms", round(dt * 1e3, 1), kc.geometry(), {k: round(v, 1) for k, v in kc.stage_ms().items() if v > 0.3}, kc.debug_counters(), file=sys.stderr)
