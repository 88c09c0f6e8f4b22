"""CPU suite for the sparse-DP chaining checker: oracle/sdp_oracle.c (this repo's C restatement of
frag_line_BCC / frag_line_remain) against (1) the committed golden outputs of the unmodified
reference (tests/golden/sdp_golden.npz: recorded whole-program call streams + synthetic sets) and
(2), where oracle/_ref/liblamsa_ref.so was built, the reference itself on fresh random inputs."""
import os

import numpy as np
import pytest

import _sdp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdp_golden.npz")


def golden_groups():
    d = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in d.files})
    return d, names


def load_group(d, name):
    rs = _sdp.ReadSet.from_dict(d, name + "/")
    return rs, (d[name + "/s1"], d[name + "/o1"]), (d[name + "/s2"], d[name + "/o2"])


def test_golden_has_recorded_and_synthetic_groups():
    _, names = golden_groups()
    assert any(n.startswith("rec_") for n in names) and any(n.startswith("syn_") for n in names)


@pytest.mark.parametrize("name", golden_groups()[1])
def test_oracle_matches_reference_golden(name):
    d, _ = golden_groups()
    rs, e1, e2 = load_group(d, name)
    o1, o2, pairs = _sdp.oracle_run(rs)
    bad = _sdp.diff_streams(o1, e1, name + " stage1") + _sdp.diff_streams(o2, e2, name + " stage2")
    assert not bad, "\n".join(bad)
    assert pairs[0] > 0


def test_oracle_empty_and_degenerate_reads():
    """reads without seeds / with one seed / one hit, and an empty batch"""
    para = _sdp.default_para()
    reads = np.array([(0, 20, 2050, 0, 0, 0, 0), (1, 20, 2050, 0, 0, 0, 0), (2, 3, 300, 0, 1, 1, 0)], dtype=_sdp.READ_DTYPE)
    hits = np.array([(1000, 0, 0, 0, 1), (5000, 0, 1, 0, 1), (5100, 0, 0, 0, 1)], dtype=_sdp.HIT_DTYPE)
    rs = _sdp.ReadSet(para, reads, [3, 1, 2], [1, 1, 1], hits)
    o1, o2, _ = _sdp.oracle_run(rs)
    assert o1[0][o1[1][0]] == 0                     # no seeds -> no skeleton
    assert o1[0][o1[1][2]] == 1                     # two co-linear seeds -> one skeleton
    if _sdp.have_ref():
        r1, r2, _ = _sdp.ref_run(rs)
        assert not _sdp.diff_streams(o1, r1) and not _sdp.diff_streams(o2, r2)
    empty = rs.subset(np.zeros(0, np.int64))
    e1, e2, _ = _sdp.oracle_run(empty)
    assert len(e1[0]) == 0 and len(e2[0]) == 0


@pytest.mark.skipif(not _sdp.have_ref(), reason="oracle/_ref/liblamsa_ref.so not built (needs the reference tree)")
@pytest.mark.parametrize("mode,seed,rf,sv,miss", [("default", 1, 0.15, 0.3, 0.3), ("default", 2, 0.5, 0.6, 0.1),
                                                  ("pacbio", 3, 0.2, 0.3, 0.3), ("ont2d", 4, 0.2, 0.5, 0.2),
                                                  ("pacbio", 5, 0.02, 0.05, 0.7)])
def test_oracle_matches_reference_live(mode, seed, rf, sv, miss):
    rs = _sdp.gen_reads(120, seed=seed, mode=mode, repeat_frac=rf, sv_rate=sv, miss_frac=miss, read_len=(300, 9000))
    r1, r2, _ = _sdp.ref_run(rs)
    o1, o2, _ = _sdp.oracle_run(rs)
    bad = _sdp.diff_streams(o1, r1, "stage1") + _sdp.diff_streams(o2, r2, "stage2")
    assert not bad, "\n".join(bad)
