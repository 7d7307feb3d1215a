"""Per-rank cost of the sharded pipeline's shard-side stages (gather, per-bucket counting, table), measured on ONE
GPU with the single-process form of the group: `world` contexts share the device and run one after the other, so
each context's stage times are that rank's compute cost without NVLink or waiting.  Shows how the canonical k-mer
density (twice the average at the low end of k-mer space, near zero at the high end) skews the ranks' work.

   python tools/skew_probe.py [world=8] [reads_per_rank=7500000] [genome=100000000] [K=25]
"""
import json
import sys

sys.path.insert(0, ".")
from allpathslg_b200 import KmerCounter, KmerGroup, synth_params

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_per = int(sys.argv[2]) if len(sys.argv) > 2 else 7_500_000
G = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000_000
K = int(sys.argv[4]) if len(sys.argv) > 4 else 25
sp = synth_params(G, 100)
kcs = []
for r in range(world):
    kc = KmerCounter(K)
    kc.synth_reads(sp, r * n_per, n_per)
    kcs.append(kc)
with KmerGroup.local(kcs) as grp:
    for _ in range(3):
        grp.count()
    st = grp.stats()
    rows = []
    for r, kc in enumerate(kcs):
        ms = kc.stage_ms()
        ni, nd = kc.totals()
        rows.append({"rank": r, "instances": ni, "distinct": nd, "owner": round(ms.get("owner", 0), 3), "local": round(ms.get("local", 0), 3),
                     "table": round(ms.get("table", 0), 3), "scatter1": round(ms.get("scatter1", 0), 3)})
    print(json.dumps({"world": world, "prefix_bits": st["prefix_bits"], "split_bits": st["split_bits"], "totals": grp.totals(), "ranks": rows}))
for kc in kcs:
    kc.close()
