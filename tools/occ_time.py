"""Stage timings of the occurrence build and the bulk lookups at the bench workload (argv: reads genome K L)."""
import sys
sys.path.insert(0, ".")
import torch
from allpathslg_b200 import KmerCounter, synth_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 25
L = int(sys.argv[4]) if len(sys.argv) > 4 else 100
kc = KmerCounter(K)
kc.synth_reads(synth_params(G, L), 0, n)
kc.finish()
for _ in range(2):
    info = kc.build_occurrences()
print(K, "records", {k: round(v, 1) for k, v in info["ms"].items()}, "total", round(sum(info["ms"].values()), 1))
if len(sys.argv) > 5:
    tb, _ = kc.read_store_info()
    out = torch.empty(tb, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ms = kc.read_freqs_device(out.data_ptr())
    print(K, "lookups", {k: round(v, 1) for k, v in ms.items()})
