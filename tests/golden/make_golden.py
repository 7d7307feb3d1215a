"""Writes tests/golden/known_answers.json.

PARITY UNPINNED: the reference source was not available, so these are not
outputs of the reference.  They are (a) HAND-COMPUTED answers for tiny inputs, written out below
as literals and only *checked* by this script, and (b) spectra of seeded synthetic read sets on
which the two independently written oracles (oracle/kmer_oracle.c and oracle/oracle_b.py) agree;
the script refuses to write a vector the two oracles disagree on.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle_a as A  # noqa: E402
from oracle import oracle_b as B  # noqa: E402

# ---- (a) hand-computed.  k-mer integers: A=0 C=1 G=2 T=3, first base most significant.
HAND = [
    # ACGT, K=2: windows AC, CG, GT.  rc(AC)=GT so AC and GT share canonical AC=0b0001=1 (count 2);
    # CG is its own reverse complement (palindrome) = 0b0110 = 6, counted once per instance.
    dict(reads=["ACGT"], K=2, kmers=[1, 6], counts=[2, 1], spectrum=[0, 1, 1]),
    # AAAA, K=2: AA x3; rc(AA)=TT > AA so canonical AA=0.
    dict(reads=["AAAA"], K=2, kmers=[0], counts=[3], spectrum=[0, 0, 0, 1]),
    # TTTT, K=2: TT x3 -> canonical AA=0 as well.
    dict(reads=["TTTT"], K=2, kmers=[0], counts=[3], spectrum=[0, 0, 0, 1]),
    # two reads, one shorter than K (yields nothing), K=3: ACG -> rc CGT; canonical ACG = 0b000110 = 6.
    dict(reads=["AC", "ACG"], K=3, kmers=[6], counts=[1], spectrum=[0, 1]),
    # a read and its reverse complement give identical canonical k-mers: GATTACA / TGTAATC, K=4.
    # GATTACA: GATT(rc AATC) ATTA(rc TAAT) TTAC(rc GTAA) TACA(rc TGTA)
    #   canonical: AATC, ATTA, GTAA, TACA ; each appears twice over the two reads.
    # AATC=0b00001101=13  ATTA=0b00111100=60  GTAA=0b10110000=176  TACA=0b11000100=196
    dict(reads=["GATTACA", "TGTAATC"], K=4, kmers=[13, 60, 176, 196], counts=[2, 2, 2, 2], spectrum=[0, 0, 4]),
    # K=1: bases A,C,G,T fold to A (A/T) and C (C/G): ACGTT -> A:3 (A,T,T) C:2 (C,G)
    dict(reads=["ACGTT"], K=1, kmers=[0, 1], counts=[3, 2], spectrum=[0, 0, 1, 1]),
    # empty input
    dict(reads=[], K=5, kmers=[], counts=[], spectrum=[0]),
]


# ---- (a2) hand-computed OCCURRENCE records: per k-mer (table order) the list of (read id, signed position);
# position 1-based, negative when the canonical k-mer is the reverse complement of the read's window.
OCC_HAND = [
    # ACGT K=2: AC forward at 1; GT at 3 is AC read backwards; the palindrome CG counts as forward.
    dict(reads=["ACGT"], K=2, kmers=[1, 6], occ=[[[0, 1], [0, -3]], [[0, 2]]]),
    # GATTACA: GATT->AATC(rc) ATTA(fw) TTAC->GTAA(rc) TACA(fw); TGTAATC: TGTA->TACA(rc) GTAA(fw) TAAT->ATTA(rc) AATC(fw)
    dict(reads=["GATTACA", "TGTAATC"], K=4, kmers=[13, 60, 176, 196],
         occ=[[[0, -1], [1, 4]], [[0, 2], [1, -3]], [[0, -3], [1, 2]], [[0, 4], [1, -1]]]),
    # even K, palindromes count as forward: ACGTACGT K=4 -> ACGT(+1, pal) CGTA(+2) GTAC(+3, pal) TACG->CGTA(-4) ACGT(+5)
    # ACGT=0b00011011=27  CGTA=0b01101100=108  GTAC=0b10110001=177
    dict(reads=["ACGTACGT"], K=4, kmers=[27, 108, 177], occ=[[[0, 1], [0, 5]], [[0, 2], [0, -4]], [[0, 3]]]),
    # a read shorter than K still owns read id 0
    dict(reads=["AC", "ACG"], K=3, kmers=[6], occ=[[[1, 1]]]),
    # reads without bases keep their ids; TT is AA read backwards
    dict(reads=["", "AAAA", "", "TTTT"], K=2, kmers=[0], occ=[[[1, 1], [1, 2], [1, 3], [3, -1], [3, -2], [3, -3]]]),
]


def occ_digest(run_off, rid, pos):
    import hashlib

    h = hashlib.sha256()
    h.update(np.ascontiguousarray(run_off, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(rid, dtype=np.uint32).tobytes())
    h.update(np.ascontiguousarray(pos, dtype=np.int32).tobytes())
    return h.hexdigest()


def occ_from_b(reads, K):
    """oracle B's occurrences flattened to oracle A's array form."""
    ob = B.occurrences(reads, K)
    ro = np.zeros(len(ob) + 1, dtype=np.uint64)
    rid, pos = [], []
    for i, (_, lst) in enumerate(ob):
        ro[i + 1] = ro[i] + np.uint64(len(lst))
        rid += [r for r, _ in lst]
        pos += [q for _, q in lst]
    return [k for k, _ in ob], ro, np.array(rid, dtype=np.uint32), np.array(pos, dtype=np.int32)


def check_occ_hand(v):
    kb, ro, rid, pos = occ_from_b(v["reads"], v["K"])
    assert kb == v["kmers"], (v, kb)
    flat = [x for lst in v["occ"] for x in lst]
    assert [int(x) for x in ro] == [0] + list(np.cumsum([len(l) for l in v["occ"]])), (v, ro)
    assert [[int(a), int(b)] for a, b in zip(rid, pos)] == flat, (v, rid, pos)
    p, o = A.pack_strings(v["reads"])
    k, c, n = A.count(p, o, v["K"])
    aro, arid, apos = A.occurrences(p, o, v["K"], k, n)
    assert (aro == ro).all() and (arid == rid).all() and (apos == pos).all(), (v, aro, arid, apos)


def occ_synth_vector(s):
    sp = A.synth_params(s["genome_len"], s["read_len"])
    p, o = A.synth_reads(sp, 0, s["n_reads"])
    k, c, n = A.count(p, o, s["K"])
    aro, arid, apos = A.occurrences(p, o, s["K"], k, n)
    _, ro, rid, pos = occ_from_b(B.unpack_reads(p, o), s["K"])
    assert (aro == ro).all() and (arid == rid).all() and (apos == pos).all(), "oracles disagree on occurrences"
    out = dict(s)
    out.update(n_instances=int(n), n_distinct=len(k), occ_sha256=occ_digest(aro, arid, apos),
               first_occ=[[int(a), int(b)] for a, b in zip(arid[:4], apos[:4])])
    return out


def check_hand(v):
    pairs = B.count_reads(v["reads"], v["K"])
    assert [k for k, _ in pairs] == v["kmers"], (v, pairs)
    assert [c for _, c in pairs] == v["counts"], (v, pairs)
    assert B.spectrum(pairs) == v["spectrum"], (v, B.spectrum(pairs))
    p, o = A.pack_strings(v["reads"])
    k, c, _ = A.count(p, o, v["K"])
    assert [int(x) for x in k[:, -1]] == v["kmers"] and [int(x) for x in c] == v["counts"]


# ---- (b) seeded synthetic sets, both oracles must agree
SYNTH = [dict(genome_len=20000, read_len=100, n_reads=6000, K=25), dict(genome_len=20000, read_len=100, n_reads=6000, K=24),
         dict(genome_len=8000, read_len=150, n_reads=1500, K=48), dict(genome_len=8000, read_len=250, n_reads=800, K=96),
         dict(genome_len=12000, read_len=100, n_reads=3000, K=20), dict(genome_len=9000, read_len=120, n_reads=1500, K=64)]


def synth_vector(s):
    sp = A.synth_params(s["genome_len"], s["read_len"])
    p, o = A.synth_reads(sp, 0, s["n_reads"])
    k, c, n = A.count(p, o, s["K"])
    reads = B.unpack_reads(p, o)
    pairs = B.count_reads(reads, s["K"])
    W = A.n_words(s["K"])
    ka = [sum(int(k[i, j]) << (64 * (W - 1 - j)) for j in range(W)) for i in range(len(k))]
    assert ka == [x for x, _ in pairs], "oracles disagree on k-mers"
    assert [int(x) for x in c] == [y for _, y in pairs], "oracles disagree on counts"
    spec = A.spectrum(c)
    # checksum of the (kmer, count) table: sum over records of (kmer mod 2^61-1) * count mod 2^61-1
    M = (1 << 61) - 1
    chk = 0
    for x, y in pairs:
        chk = (chk + (x % M) * y) % M
    out = dict(s)
    out.update(n_instances=int(n), n_distinct=len(pairs), spectrum=[int(x) for x in spec], table_checksum=chk,
               first_kmers=[str(x) for x in ka[:4]], first_counts=[int(x) for x in c[:4]])
    return out


def main():
    for v in HAND:
        check_hand(v)
    for v in OCC_HAND:
        check_occ_hand(v)
    out = dict(note="parity unpinned: hand-computed + two-oracle-agreement vectors, not reference outputs",
               hand=HAND, synth=[synth_vector(s) for s in SYNTH], occ_hand=OCC_HAND,
               occ_synth=[occ_synth_vector(s) for s in (SYNTH[0], SYNTH[2], SYNTH[3])])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
