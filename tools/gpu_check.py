"""Development aid: run the CUDA path against the oracle on a ladder of cases and print diagnostics."""
import sys, time, random, traceback
import numpy as np
sys.path.insert(0, ".")
from allpathslg_b200 import KmerCounter, synth_params
from oracle import oracle_a as A

def compare(name, packed, off, K, prefix_bits=0, uniform=None):
    t = time.time()
    ek, ec, en = A.count(packed, off, K)
    es = A.spectrum(ec)
    kc = KmerCounter(K, prefix_bits=prefix_bits)
    try:
        if uniform: kc.add_reads_uniform(packed, uniform[0], uniform[1])
        else: kc.add_reads(packed, off)
        kc.finish()
        ni, nd = kc.totals()
        geo = kc.geometry()
        ok = True
        msg = []
        if ni != en: ok = False; msg.append("instances %d != %d" % (ni, en))
        if nd != len(ek): ok = False; msg.append("distinct %d != %d" % (nd, len(ek)))
        gs = kc.spectrum()
        if len(gs) != len(es) or not (gs == es).all():
            ok = False; msg.append("spectrum differs: got[:6]=%s exp[:6]=%s len %d/%d" % (gs[:6], es[:6], len(gs), len(es)))
        gk, gc = kc.counts()
        if nd == len(ek):
            if not (gk == ek).all():
                bad = np.nonzero((gk != ek).any(axis=1))[0]
                ok = False; msg.append("kmers differ at %d rows, first %d: got %s exp %s" % (len(bad), bad[0], gk[bad[0]], ek[bad[0]]))
            if not (gc.astype(np.uint64) == ec).all():
                bad = np.nonzero(gc.astype(np.uint64) != ec)[0]
                ok = False; msg.append("counts differ at %d rows first %d" % (len(bad), bad[0]))
        # lookups
        if nd and ok:
            idx = np.random.RandomState(1).randint(0, nd, size=min(nd, 1000))
            q = ek[idx]
            r = kc.lookup(q, canonicalise=False)
            if not (r.astype(np.uint64) == ec[idx]).all(): ok = False; msg.append("lookup mismatch")
            rf = kc.read_freqs()
            erf = A.read_freqs(packed, off, K, ek, ec)
            erf32 = np.where(erf == np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0xFFFFFFFF), erf).astype(np.uint32)
            if not (rf == erf32).all(): ok = False; msg.append("read_freqs mismatch at %d" % np.nonzero(rf != erf32)[0][0])
        print("%-34s K=%-3d %s n=%d nd=%d geo=%s %.2fs %s" % (name, K, "OK  " if ok else "FAIL", en, len(ek), geo, time.time() - t, "; ".join(msg)))
        return ok
    except Exception as e:
        print("%-34s K=%-3d EXC %r" % (name, K, e)); traceback.print_exc()
        return False
    finally:
        kc.close()

def main():
    random.seed(7)
    allok = True
    p, o = A.pack_strings(["ACGT"]); allok &= compare("ACGT", p, o, 2)
    for K in [1, 3, 11, 25, 32, 33, 64, 65, 96]:
        reads = ["".join(random.choice("ACGT") for _ in range(random.choice([0, 1, K - 1, K, K + 1, K + 7, 150, 100]))) for _ in range(300)]
        reads += ["A" * (K + 300), "ACGT" * 100, reads[3], "T" * (K + 20)]
        p, o = A.pack_strings(reads)
        allok &= compare("ragged-random", p, o, K)
        allok &= compare("ragged-random P=12", p, o, K, prefix_bits=12)
    # poly-A heavy: oversize buckets
    reads = ["A" * 200 for _ in range(400)] + ["ACGTACGTAC" * 20 for _ in range(300)] + ["".join(random.choice("AC") for _ in range(120)) for _ in range(500)]
    p, o = A.pack_strings(reads)
    for K in [5, 25, 40, 96]:
        allok &= compare("low-complexity(big buckets)", p, o, K)
        allok &= compare("low-complexity P=4", p, o, K, prefix_bits=4)
    # synthetic genome configs, growing
    for (G, n, L, K) in [(50_000, 20_000, 100, 25), (1_000_000, 400_000, 100, 25), (1_000_000, 200_000, 100, 24),
                         (500_000, 60_000, 250, 48), (500_000, 60_000, 250, 64), (500_000, 60_000, 250, 96), (500_000, 100_000, 100, 20)]:
        sp = A.synth_params(G, L)
        p, o = A.synth_reads(sp, 0, n)
        allok &= compare("synth G=%d n=%d L=%d" % (G, n, L), p, o, K, uniform=(n, L))
    # E. coli size
    sp = A.synth_params(4_600_000, 100)
    p, o = A.synth_reads(sp, 0, 2_300_000)
    allok &= compare("ecoli-size", p, o, 25, uniform=(2_300_000, 100))
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1

if __name__ == "__main__":
    sys.exit(main())
