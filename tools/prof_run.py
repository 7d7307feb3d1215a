"""Small driver for ncu: one pipeline pass over a synthetic read set (size via argv)."""
import sys
sys.path.insert(0, ".")
from allpathslg_b200 import KmerCounter, synth_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else n * 100 // 60
K = int(sys.argv[3]) if len(sys.argv) > 3 else 25
L = int(sys.argv[4]) if len(sys.argv) > 4 else 100
kc = KmerCounter(K)
kc.synth_reads(synth_params(G, L), 0, n)
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
for _ in range(reps):
    kc.finish()
print(kc.totals(), kc.geometry(), {k: round(v, 3) for k, v in kc.stage_ms().items()})
