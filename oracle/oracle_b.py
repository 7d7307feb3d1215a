"""oracle/oracle_b.py -- CPU oracle "B": an independently written restatement.

TEST INFRASTRUCTURE, NOT PRODUCT.  PARITY UNPINNED (no reference source was
available; see the header of kmer_oracle.c and SURVEY.md section 0).

Oracle B shares no code and no representation with oracle A: it works on
ASCII strings, canonicalises by *string* comparison of a k-mer with its
reverse-complement string (equal to the integer rule because A<C<G<T and the
first base is most significant), and counts with a dict.  Pure-Python loops:
small cases only.
"""
from collections import Counter

_COMP = str.maketrans("ACGT", "TGCA")


def unpack_reads(packed, off):
    """2-bit LE packed bytes + base offsets -> list of ACGT strings."""
    reads = []
    for r in range(len(off) - 1):
        s = []
        for q in range(int(off[r]), int(off[r + 1])):
            s.append("ACGT"[(int(packed[q >> 2]) >> ((q & 3) * 2)) & 3])
        reads.append("".join(s))
    return reads


def canonical_str(kmer):
    rc = kmer.translate(_COMP)[::-1]
    return kmer if kmer <= rc else rc


def kmer_to_int(kmer):
    v = 0
    for ch in kmer:
        v = (v << 2) | "ACGT".index(ch)
    return v


def count_reads(reads, K):
    """-> (sorted list of (kmer_int, count)), using string canonicalisation."""
    c = Counter()
    for s in reads:
        for p in range(0, len(s) - K + 1):
            c[canonical_str(s[p:p + K])] += 1
    return sorted((kmer_to_int(k), n) for k, n in c.items())


def spectrum(pairs):
    """dense spectrum list, index = frequency."""
    if not pairs:
        return [0]
    mx = max(n for _, n in pairs)
    s = [0] * (mx + 1)
    for _, n in pairs:
        s[n] += 1
    return s


def occurrences(reads, K):
    """-> sorted list of (kmer_int, [(read_id, signed_pos), ...]): every instance of every canonical k-mer,
    instances in (read id, position) order; signed_pos is 1-based, negative when the canonical string is the
    reverse complement of the read's window (a palindrome is forward)."""
    d = {}
    for r, s in enumerate(reads):
        for p in range(0, len(s) - K + 1):
            w = s[p:p + K]
            c = canonical_str(w)
            d.setdefault(c, []).append((r, (p + 1) if c == w else -(p + 1)))
    return sorted((kmer_to_int(k), v) for k, v in d.items())
