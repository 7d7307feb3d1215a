"""The hook that pins parity the day the reference source is mounted: `make -C oracle _ref` compiles the reference's own
k-mer layer from /root/reference (never copied into this repo).  With the empty mount of rounds 1-2 the target is a
no-op and the pinning test skips -- the oracle stays a spec-derived restatement ("parity unpinned")."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_ref_target_runs_and_reports():
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "_ref"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    have_src = bool(glob.glob(os.path.join(REF, "src", "kmers", "KmerSpectra.cc")))
    if have_src:
        assert os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_kmers.so")), out.stdout
    else:
        assert "parity unpinned" in out.stdout


def test_oracle_against_the_reference_build():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_kmers.so")
    if not os.path.exists(so):
        pytest.skip("no reference build (the reference mount holds no source): parity unpinned")
    # The day this runs: drive the reference's KmerSpectrum / SortKmers entry points on the synthetic configs of
    # SURVEY.md section 8(d) and diff them with oracle_a.count / oracle_a.spectrum; until the signatures can be read
    # (Appendix A.2, questions 1-9) nothing can be asserted here.
    pytest.fail("a reference build exists: write the pinning comparison against its entry points (SURVEY.md Appendix A)")
