"""GPU suite (B200): record statistics over the resident 2-bit reference (lb2_aux_run, aux_scan.cuh -- the
reference-touching half of lamsa_res_aux, src/frag_check.c:793-853) against a numpy walk of the same CIGARs."""
import numpy as np
import pytest

import lamsa_b200

pytestmark = pytest.mark.gpu
M, I, D, S, H = 0, 1, 2, 4, 5


def walk(cigar, read, ref, pac0):
    """The loop of src/frag_check.c:810-832 on unpacked bases."""
    ri = fi = nm = nmm = nio = nie = ndo = nde = 0
    for w in cigar:
        op, ln = int(w) & 15, int(w) >> 4
        if op == M:
            a, b = read[ri:ri + ln], ref[pac0 + fi: pac0 + fi + ln]
            mm = int((a != b).sum())
            nmm += mm; nm += ln - mm; ri += ln; fi += ln
        elif op == I:
            ri += ln; nie += ln; nio += 1
        elif op == D:
            fi += ln; nde += ln; ndo += 1
        elif op == S:
            ri += ln
        else:
            return [nm, nmm, nio, nie, ndo, nde, ri, -(op + 1)]
    return [nm, nmm, nio, nie, ndo, nde, ri, fi]


def test_aux_counts_match_a_cigar_walk():
    rng = np.random.default_rng(5)
    L = 400_000
    ref = rng.integers(0, 4, size=L, dtype=np.uint8)
    q4 = ref.reshape(-1, 4)
    pac = (q4[:, 0] << 6 | q4[:, 1] << 4 | q4[:, 2] << 2 | q4[:, 3]).astype(np.uint8)      # MSB-first, src/bntseq.c:242
    ctx = lamsa_b200.Context(0)
    ctx.set_reference(np.concatenate((pac, np.zeros(8, np.uint8))), L)
    cigars, reads, pacs = [], [], []
    for k in range(3000):
        pos = int(rng.integers(0, L - 30000))
        nops = int(rng.integers(1, 60))
        cg, rd, fi = [], [], pos
        if rng.random() < 0.3:
            s = int(rng.integers(1, 300)); cg.append(s << 4 | S); rd.append(rng.integers(0, 4, size=s, dtype=np.uint8))
        for _ in range(nops):
            op = int(rng.choice([M, M, M, I, D]))
            ln = int(rng.integers(1, 400 if op == M else 40))
            if cg and (cg[-1] & 15) == op:
                continue
            cg.append(ln << 4 | op)
            if op == M:
                seg = ref[fi:fi + ln].copy()
                flip = rng.random(ln) < 0.08
                seg[flip] = (seg[flip] + rng.integers(1, 4, size=int(flip.sum()))) & 3
                if rng.random() < 0.05:
                    seg[rng.integers(0, ln)] = 4                          # N in the read
                rd.append(seg); fi += ln
            elif op == I:
                rd.append(rng.integers(0, 4, size=ln, dtype=np.uint8))
            else:
                fi += ln
        if k % 97 == 0:
            cg.append(7 << 4 | H)                                         # unexpected operator: reported, not walked
        if rng.random() < 0.3:
            s = int(rng.integers(1, 300)); cg.append(s << 4 | S); rd.append(rng.integers(0, 4, size=s, dtype=np.uint8))
        cigars.append(np.array(cg, dtype=np.int32)); reads.append(np.concatenate(rd) if rd else np.zeros(0, np.uint8)); pacs.append(pos)
    cigars.append(np.zeros(0, np.int32)); reads.append(np.zeros(0, np.uint8)); pacs.append(0)      # empty record
    got = ctx.aux_counts(cigars, reads, pacs)
    exp = np.array([walk(c, r, ref, p) for c, r, p in zip(cigars, reads, pacs)], dtype=np.int32)
    bad_op = exp[:, 7] < 0
    assert (got[bad_op, 7] == exp[bad_op, 7]).all()
    assert (got[~bad_op] == exp[~bad_op]).all(), np.nonzero((got != exp).any(axis=1) & ~bad_op)[0][:5]
    with pytest.raises(RuntimeError):          # needs a resident reference
        lamsa_b200.Context(0).aux_counts(cigars[:1], reads[:1], pacs[:1])
    ctx.close()
