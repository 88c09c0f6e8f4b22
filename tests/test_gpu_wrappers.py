"""GPU suite (B200): the drop-in symbols ksw_extend_c / ksw_extend_r / sw_mid_fix / ksw_bi_extend of
liblamsa_b200.so, called directly through the C ABI, against the oracle's restatements (which
tests/test_wrappers_oracle.py pins against the unmodified reference src/ksw.c:809-926) on >= 10 000 seeded
pairs covering every exit of ksw_bi_extend incl. the aln_mode & 2 float comparison (:881,:900)."""
import collections

import pytest

import _wrappers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n", [(911, 5000), (912, 5000), (913, 600)])
def test_gpu_wrappers_match_oracle(ctx, seed, n):
    pairs = _wrappers.gen_pairs(n, seed, qmax=500 if seed != 913 else 3000)
    gpu, orc = _wrappers.Impl("gpu"), _wrappers.Impl("oracle")
    a, b = _wrappers.run_all(gpu, pairs), _wrappers.run_all(orc, pairs)
    exits = collections.Counter()
    for k, (ra, rb, p) in enumerate(zip(a, b, pairs)):
        assert ra == rb, f"pair {k} (qlen {len(p[0])}, tlen {len(p[1])}): GPU {ra} != oracle {rb}"
        exits[_wrappers.classify_exit(rb[0], rb[1], p[0], p[1], p[4])] += 1
    for e in ("left_end", "left_global", "right_end", "right_global", "mid_clip", "mid_global"):
        assert exits[e] >= (5 if n >= 5000 else 1), exits
