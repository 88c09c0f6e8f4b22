"""CPU suite: the oracle's restatements of the host wrappers -- orc_extend_c / orc_extend_r /
orc_mid_fix / orc_bi_extend (oracle/dp_oracle.c) -- pinned against the unmodified reference's
ksw_extend_c / ksw_extend_r / sw_mid_fix / ksw_bi_extend (src/ksw.c:809-926, oracle/_ref/libksw_ref.so)
on seeded pairs that reach every exit of ksw_bi_extend, including the float comparison of :881 with
aln_mode & 2.  Only possible where the reference was compiled (this container)."""
import collections

import pytest

import _oracle
import _wrappers

pytestmark = pytest.mark.skipif(not _oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("seed,n", [(901, 6000), (902, 6000)])
def test_oracle_wrappers_match_reference(seed, n):
    pairs = _wrappers.gen_pairs(n, seed)
    ref, orc = _wrappers.Impl("ref"), _wrappers.Impl("oracle")
    a, b = _wrappers.run_all(ref, pairs), _wrappers.run_all(orc, pairs)
    exits = collections.Counter()
    for k, (ra, rb, p) in enumerate(zip(a, b, pairs)):
        assert ra == rb, f"pair {k} (qlen {len(p[0])}, tlen {len(p[1])}): reference {ra} != oracle {rb}"
        exits[_wrappers.classify_exit(ra[0], ra[1], p[0], p[1], p[4])] += 1
    # the generator must really reach all six exits (five of ksw_bi_extend, sw_mid_fix's two branches)
    for e in ("left_end", "left_global", "right_end", "right_global", "mid_clip", "mid_global"):
        assert exits[e] >= 10, exits
