"""Small driver for ncu: one pipeline pass + one occurrence-record build over a synthetic read set."""
import sys
sys.path.insert(0, ".")
from allpathslg_b200 import KmerCounter, synth_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else n * 100 // 60
K = int(sys.argv[3]) if len(sys.argv) > 3 else 25
L = int(sys.argv[4]) if len(sys.argv) > 4 else 100
kc = KmerCounter(K)
kc.synth_reads(synth_params(G, L), 0, n)
kc.finish()
print(kc.totals(), kc.build_occurrences())
