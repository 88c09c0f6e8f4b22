"""The C restatement of the seed-and-chain half of hash_split_map (oracle/hash_oracle.c) against the unmodified
reference functions (init_hash, hash_hit, hash_main_line of src/split_mapping.c through oracle/hash_ref_shim.c):
identical lines (read position, diagonal, relation of every node) on seeded SV-shaped inputs.  Runs where the
reference was compiled (this container); elsewhere the committed golden lines pin the oracle."""
import os

import numpy as np
import pytest

import _hash

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hash_line_golden.npz")


def test_hash_oracle_matches_golden_lines():
    g = np.load(GOLDEN)
    cases = _hash.gen_cases(int(g["n"]), int(g["seed"]))
    off = g["off"]
    for k, c in enumerate(cases):
        got = _hash.oracle_line(c)
        want = g["lines"][off[k]:off[k + 1]]
        assert got.shape == want.shape and (got == want).all(), f"case {k}: oracle line differs from the reference's"
    assert sum(len(c["read"]) for c in cases) == int(g["read_bases"])


@pytest.mark.skipif(not _hash.have_ref(), reason="oracle/_ref/liblamsa_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_hash_oracle_matches_reference(seed):
    cases = _hash.gen_cases(500, seed)
    n_nodes = 0
    for k, c in enumerate(cases):
        got, want = _hash.oracle_line(c), _hash.ref_line(c)
        assert got.shape == want.shape and (got == want).all(), f"seed {seed} case {k}: {got.tolist()[:8]} vs {want.tolist()[:8]}"
        n_nodes += len(want)
    assert n_nodes > 5000
