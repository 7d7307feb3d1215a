import sys, random, os
sys.path.insert(0, ".")
os.environ["APGK_DEBUG"] = "1"
import numpy as np
from allpathslg_b200 import KmerCounter
from oracle import oracle_a as A
random.seed(7)
for K in [1, 3, 11, 25, 32, 33, 64]:
    reads = ["".join(random.choice("ACGT") for _ in range(random.choice([0, 1, K - 1, K, K + 1, K + 7, 150, 100]))) for _ in range(300)]
    reads += ["A" * (K + 300), "ACGT" * 100, reads[3], "T" * (K + 20)]
p, o = A.pack_strings(reads)
ek, ec, en = A.count(p, o, 64)
for rep in range(3):
  for P in [12, 10, 8, 6, 0]:
    kc = KmerCounter(64, prefix_bits=P)
    kc.add_reads(p, o); kc.finish()
    print("P", P, kc.totals(), (en, len(ek)), kc.geometry(), flush=True)
    kc.close()
