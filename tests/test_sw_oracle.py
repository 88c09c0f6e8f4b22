"""The scalar restatement of the striped local Smith-Waterman (oracle/sw_oracle.c) against the unmodified reference
(ksw_align2 of src/ksw.c through oracle/_ref/libksw_ref.so): all seven kswr_t fields on seeded pairs, both profile
widths, every flag, byte overflow."""
import os

import numpy as np
import pytest

import _sw

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sw_golden.npz")


def test_sw_oracle_matches_golden():
    g = np.load(GOLDEN)
    cases = _sw.gen_cases(int(g["n"]), int(g["seed"]))
    for k, c in enumerate(cases):
        assert _sw.oracle_align2(c) == tuple(int(v) for v in g["res"][k]), f"case {k}"


@pytest.mark.skipif(not _sw.have_ref(), reason="oracle/_ref/libksw_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [31, 32])
def test_sw_oracle_matches_reference(seed):
    cases = _sw.gen_cases(1500, seed)
    seen255 = seen2 = 0
    for k, c in enumerate(cases):
        got, want = _sw.oracle_align2(c), _sw.ref_align2(c)
        assert got == want, f"seed {seed} case {k} (xtra {c['xtra']:#x}, qlen {len(c['q'])}, tlen {len(c['t'])}): {got} vs {want}"
        seen255 += want[0] == 255; seen2 += want[3] > 0
    assert seen255 > 5 and seen2 > 50
