"""GPU suite: sparse-DP chaining (lb2_sdp_*) against the golden outputs of the unmodified reference
(tests/golden/sdp_golden.npz) and against the oracle restatement on fresh random seed-hit sets.
Bit-exact: the skeleton streams (lines, fragments, seeds) must be identical word for word."""
import os

import numpy as np
import pytest

import _sdp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdp_golden.npz")


@pytest.fixture(scope="module")
def ctx():
    import lamsa_b200
    c = lamsa_b200.Context(0)
    yield c
    c.close()


def gpu_run(ctx, rs):
    from lamsa_b200.sdp import SdpBatch
    b = SdpBatch(ctx, rs.para, rs.reads, rs.seed_id, rs.map_n, rs.hits)
    s1 = b.run_bcc()
    p1 = b.stats()["pairs"]
    s2 = b.run_remain(rs.reads, rs.regs)
    p2 = b.stats()["pairs"]
    b.close()
    return s1, s2, (p1, p2)


def group_names():
    d = np.load(GOLDEN)
    return sorted({k.split("/")[0] for k in d.files})


@pytest.mark.parametrize("name", group_names())
def test_gpu_sdp_matches_reference_golden(ctx, name):
    d = np.load(GOLDEN)
    rs = _sdp.ReadSet.from_dict(d, name + "/")
    e1, e2 = (d[name + "/s1"], d[name + "/o1"]), (d[name + "/s2"], d[name + "/o2"])
    g1, g2, _ = gpu_run(ctx, rs)
    bad = _sdp.diff_streams(g1, e1, name + " stage1") + _sdp.diff_streams(g2, e2, name + " stage2")
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("mode,seed,rf,sv,miss,n", [("default", 11, 0.15, 0.3, 0.3, 2000), ("default", 12, 0.6, 0.6, 0.1, 600),
                                                    ("pacbio", 13, 0.2, 0.3, 0.3, 600), ("ont2d", 14, 0.2, 0.5, 0.2, 600),
                                                    ("pacbio", 15, 0.02, 0.05, 0.7, 2000)])
def test_gpu_sdp_matches_oracle_random(ctx, mode, seed, rf, sv, miss, n):
    rs = _sdp.gen_reads(n, seed=seed, mode=mode, repeat_frac=rf, sv_rate=sv, miss_frac=miss, read_len=(300, 12000))
    o1, o2, opairs = _sdp.oracle_run(rs)
    g1, g2, gpairs = gpu_run(ctx, rs)
    bad = _sdp.diff_streams(g1, o1, "stage1") + _sdp.diff_streams(g2, o2, "stage2")
    assert not bad, "\n".join(bad)
    assert tuple(int(v) for v in opairs) == gpairs        # same number of edge classifications in the scans


def test_gpu_sdp_repeat_heavy_reads(ctx):
    """up to per_aln_m = 200 hits per seed on most seeds (thousands of nodes per read): scratch bounds,
    long predecessor scans, big son lists, many skeletons per cluster"""
    rs = _sdp.gen_reads(25, seed=50, mode="default", repeat_frac=0.9, sv_rate=0.5, miss_frac=0.1, max_hits=200, read_len=(2000, 9000))
    o1, o2, opairs = _sdp.oracle_run(rs)
    g1, g2, gpairs = gpu_run(ctx, rs)
    bad = _sdp.diff_streams(g1, o1, "stage1") + _sdp.diff_streams(g2, o2, "stage2")
    assert not bad, "\n".join(bad)
    assert tuple(int(v) for v in opairs) == gpairs


def test_gpu_sdp_degenerate(ctx):
    para = _sdp.default_para()
    reads = np.array([(0, 20, 2050, 0, 0, 0, 0), (1, 20, 2050, 0, 0, 0, 0), (2, 3, 300, 0, 1, 1, 0)], dtype=_sdp.READ_DTYPE)
    hits = np.array([(1000, 0, 0, 0, 1), (5000, 0, 1, 0, 1), (5100, 0, 0, 0, 1)], dtype=_sdp.HIT_DTYPE)
    rs = _sdp.ReadSet(para, reads, [3, 1, 2], [1, 1, 1], hits)
    o1, o2, _ = _sdp.oracle_run(rs)
    g1, g2, _ = gpu_run(ctx, rs)
    assert not _sdp.diff_streams(g1, o1) and not _sdp.diff_streams(g2, o2)
    empty = rs.subset(np.zeros(0, np.int64))
    e1, e2, _ = gpu_run(ctx, empty)
    assert len(e1[0]) == 0 and len(e2[0]) == 0


def test_gpu_sdp_stage2_on_a_regrouped_batch(ctx):
    """lb2_sdp_get_tracked / _set_tracked: stage 2 on a batch object that never ran stage 1, holding a
    permuted subset of the reads (what the batch producer does when reads reach stage 2 at different times)"""
    from lamsa_b200.sdp import SdpBatch
    rs = _sdp.gen_reads(300, seed=31, mode="pacbio", repeat_frac=0.3, sv_rate=0.5, miss_frac=0.2, read_len=(500, 9000))
    o1, o2, _ = _sdp.oracle_run(rs)
    a = SdpBatch(ctx, rs.para, rs.reads, rs.seed_id, rs.map_n, rs.hits)
    g1 = a.run_bcc()
    flags = a.get_tracked()
    a.close()
    assert not _sdp.diff_streams(g1, o1, "stage1")
    hit_first = np.concatenate(([0], np.cumsum([int(rs.map_n[int(r["seed_first"]):int(r["seed_first"]) + int(r["seed_out"])].sum())
                                                for r in rs.reads])))
    pick = np.random.default_rng(2).permutation(len(rs))[:173]
    sub = rs.subset(pick)
    sub_flags = np.concatenate([flags[hit_first[i]:hit_first[i + 1]] for i in pick])
    b = SdpBatch(ctx, sub.para, sub.reads, sub.seed_id, sub.map_n, sub.hits)
    b.set_tracked(sub_flags)
    s2, f2 = b.run_remain(sub.reads, sub.regs)
    b.close()
    for k, i in enumerate(pick):
        assert np.array_equal(s2[f2[k]:f2[k + 1]], o2[0][o2[1][i]:o2[1][i + 1]]), (k, i)
