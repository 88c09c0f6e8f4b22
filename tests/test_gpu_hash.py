"""GPU suite for the local split mapping (include/lamsa_b200.h section 6):
  * hash_line_kernel (k-mer index of the window, look-up, chaining) against the oracle, which is pinned against the
    unmodified reference (tests/test_hash_oracle.py): identical lines, batched;
  * the drop-in hash_split_map (GPU line + GPU DP stitching) against golden CIGARs of the unmodified reference
    (tests/golden/make_hash_split_golden.py).
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import _hash

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hash_split_golden.npz")


@pytest.mark.parametrize("seed,n", [(21, 400), (22, 400)])
def test_gpu_hash_lines_match_oracle(ctx, seed, n):
    cases = _hash.gen_cases(n, seed)
    lines, hits = _hash.gpu_lines(ctx, cases)
    nodes = 0
    for k, c in enumerate(cases):
        want = _hash.oracle_line(c)
        assert lines[k].shape == want.shape and (lines[k] == want).all(), f"seed {seed} case {k}: {lines[k].tolist()[:6]} vs {want.tolist()[:6]}"
        nodes += len(want)
    assert nodes > 3000 and sum(hits) > nodes


def test_gpu_hash_lines_small_batches_and_long_reads(ctx):
    """one request per launch (the drop-in's shape), and 12 kbp gaps (thousands of k-mers, the node pool grows)"""
    for c in _hash.gen_cases(12, 23):
        got, _ = _hash.gpu_lines(ctx, [c])
        want = _hash.oracle_line(c)
        assert got[0].shape == want.shape and (got[0] == want).all()
    big = _hash.gen_cases(10, 24, max_len=12000)
    lines, _ = _hash.gpu_lines(ctx, big)
    for k, c in enumerate(big):
        want = _hash.oracle_line(c)
        assert lines[k].shape == want.shape and (lines[k] == want).all(), f"long case {k}"


def test_gpu_hash_split_map_matches_reference_golden(ctx):
    g = np.load(GOLDEN)
    cases = _hash.gen_cases(int(g["n"]), int(g["seed"]), max_len=int(g["max_len"]))
    off = g["off"]
    flags = 0
    for k, c in enumerate(cases):
        cig, res = _hash.gpu_split_map(c)
        want = g["cigars"][off[k]:off[k + 1]]
        assert res == int(g["res"][k]), f"case {k}: return value {res} vs {int(g['res'][k])}"
        assert len(cig) == len(want) and (cig == want).all(), f"case {k}: CIGAR differs"
        flags |= res
    assert flags & 2


def test_gpu_hash_edge_cases(ctx):
    """windows / reads shorter than a k-mer, empty windows, a read of exactly one k-mer, identical sequences with a
    k-mer that repeats beyond the 50-hit cap everywhere (no seeds survive)"""
    rng = np.random.default_rng(5)
    base = dict(ref_offset=0, hash_len=10, hash_step=10, split_len=100, head=1, tail=1)
    ref = rng.integers(0, 4, size=300, dtype=np.uint8)
    cases = [dict(base, ref=ref, read=ref[:5].copy()),                         # read shorter than a k-mer: no seeds
             dict(base, ref=ref[:4].copy(), read=ref[:60].copy()),               # window shorter than a k-mer
             dict(base, ref=np.zeros(0, np.uint8), read=ref[:60].copy()),        # empty window
             dict(base, ref=ref, read=ref[40:50].copy()),                        # one k-mer
             dict(base, ref=np.zeros(400, np.uint8), read=np.zeros(200, np.uint8)),      # poly-A: every k-mer beyond the cap
             dict(base, ref=ref, read=ref.copy(), head=0, tail=0),
             dict(base, ref=ref, read=np.full(100, 4, np.uint8))]                # all-N read (hashes as G)
    lines, hits = _hash.gpu_lines(ctx, cases)
    for k, c in enumerate(cases):
        want = _hash.oracle_line(c)
        assert lines[k].shape == want.shape and (lines[k] == want).all(), f"edge case {k}"
    assert len(lines[0]) == 0 and len(lines[1]) == 0 and len(lines[2]) == 0 and hits[4] == 0 and len(lines[5]) > 20
