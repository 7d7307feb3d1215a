import sys, os, ctypes as C
sys.path.insert(0, ".")
os.environ["APGK_SYNC_DEBUG"] = "1"
import numpy as np
from allpathslg_b200 import KmerCounter, _lib
from oracle import oracle_a as A
L = _lib.lib()
sp = A.synth_params(300_000, 100)
p, o = A.synth_reads(sp, 0, 30_000)
kc = KmerCounter(25)
kc.add_reads_uniform(p, 30_000, 100)
cnts = kc.owner_plan(2)
print(cnts, flush=True)
ptr = C.c_void_p()
assert L.apgk_device_alloc(kc._h, C.byref(ptr), int(cnts.sum()) * 8 + 4096) == 0
print("lib-owned buffer", hex(ptr.value), flush=True)
try:
    kc.owner_scatter(ptr.value)
    print("owner_scatter into lib memory ok", flush=True)
except Exception as e:
    print("EXC lib mem", e, flush=True)
