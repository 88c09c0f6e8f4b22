"""Whole-program SAM fixtures at the sizes / shapes BASELINE.json's configs name, produced by running the
UNMODIFIED reference `lamsa` (scratch copy under /tmp, GEM seeding included) on seeded synthetic data.
Vectorised generator (oracle/make_sam_fixtures.py loops per base); references may have several contigs, donors
carry deletions / insertions / inversions / tandem duplications and TRANSLOCATIONS between contigs.
Test infrastructure; only possible in the container that has /root/reference.

    python oracle/make_big_fixtures.py multi      # tests/golden/sam_multi/ (committed, xz members): 4 contigs, SVs
    python oracle/make_big_fixtures.py c3full     # bigfix/sam_c3full: BASELINE configs[2] at full size
    python oracle/make_big_fixtures.py c4full     # bigfix/sam_c4full: configs[3] at full read count
    python oracle/make_big_fixtures.py c5s        # bigfix/sam_c5s: configs[4] scaled (24 contigs, 120 Mbp: what fits a gpurun snapshot; 20 k reads)
"""
import os
import shutil
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_sam_fixtures import ROOT, build_reference, compress_members  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
COMP[ACGT] = ACGT[::-1]


def write_ref(path, contigs):
    with open(path, "wb") as f:
        for name, codes in contigs:
            f.write(f">{name}\n".encode())
            s = ACGT[codes]
            n = len(s)
            full = (n // 60) * 60
            body = np.empty((full // 60, 61), dtype=np.uint8)
            body[:, :60] = s[:full].reshape(-1, 60)
            body[:, 60] = 10
            f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def make_donor(rng, contigs, sv_every, sv_max, transloc):
    """Per contig: an SV every ~sv_every bases (DEL / INS / INV / DUP, and -- when transloc -- a segment of ANOTHER
    contig spliced in), lengths 50..sv_max."""
    out = []
    for ci, (name, ref) in enumerate(contigs):
        parts, pos, n = [], 0, len(ref)
        while pos < n:
            nxt = min(n, pos + int(rng.integers(sv_every * 2 // 3, sv_every * 4 // 3)))
            parts.append(ref[pos:nxt])
            if nxt >= n:
                break
            kind = int(rng.integers(0, 5 if transloc and len(contigs) > 1 else 4))
            ln = int(rng.integers(50, sv_max))
            if kind == 0:
                nxt = min(n, nxt + ln)                                    # deletion
            elif kind == 1:
                parts.append(rng.integers(0, 4, ln, dtype=np.uint8))      # insertion
            elif kind == 2:
                seg = ref[nxt:nxt + ln]; parts.append((3 - seg[::-1]).astype(np.uint8)); nxt = min(n, nxt + ln)   # inversion
            elif kind == 3:
                parts.append(ref[max(0, nxt - ln):nxt])                   # tandem duplication
            else:                                                         # translocation from another contig
                oc = (ci + 1 + int(rng.integers(0, len(contigs) - 1))) % len(contigs)
                oref = contigs[oc][1]
                st = int(rng.integers(0, max(1, len(oref) - ln)))
                parts.append(oref[st:st + ln])
            pos = nxt
        out.append(np.concatenate(parts))
    return out


def mutate_reads(rng, segs, err, mix):
    """segs: (n_reads, L) uint8 codes -> list of mutated reads (vectorised: 1 draw per base)."""
    s, i, d = (x / sum(mix) for x in mix)
    n, L = segs.shape
    r = rng.random((n, L), dtype=np.float32)
    is_sub = r < err * s
    is_del = (~is_sub) & (r < err * (s + d))
    is_ins = (~is_sub) & (~is_del) & (r < err)
    cnt = np.ones((n, L), dtype=np.int8)
    cnt[is_del] = 0
    cnt[is_ins] = 2
    flat = segs.reshape(-1)
    cf = cnt.reshape(-1).astype(np.int64)
    start = np.cumsum(cf) - cf
    out = np.repeat(flat, cf)
    sub_pos = start[is_sub.reshape(-1)]
    out[sub_pos] = (flat[is_sub.reshape(-1)] + rng.integers(1, 4, size=sub_pos.size, dtype=np.uint8)) & 3
    ins_pos = start[is_ins.reshape(-1)]
    out[ins_pos] = rng.integers(0, 4, size=ins_pos.size, dtype=np.uint8)      # inserted base BEFORE the original one
    ends = np.cumsum(cnt.astype(np.int64).sum(axis=1))
    return out, np.concatenate(([0], ends))


def make(outdir, contig_lens, n_reads, read_len, err, seed, aln_opts=(), sv_every=0, sv_max=3000, transloc=False,
         mix=(1, 1, 1), threads=8):
    t0 = time.time()
    exe = build_reference()
    tmp = f"/tmp/lamsa_fixture_{os.path.basename(outdir)}"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    rng = np.random.default_rng(seed)
    contigs = [(f"chr{k + 1}", rng.integers(0, 4, ln, dtype=np.uint8)) for k, ln in enumerate(contig_lens)]
    write_ref(os.path.join(tmp, "ref.fa"), contigs)
    donors = make_donor(rng, contigs, sv_every, sv_max, transloc) if sv_every else [c[1] for c in contigs]
    weights = np.array([len(d) for d in donors], dtype=np.float64)
    which = rng.choice(len(donors), size=n_reads, p=weights / weights.sum())
    with open(os.path.join(tmp, "reads.fa"), "wb") as f:
        for lo in range(0, n_reads, 2000):
            hi = min(n_reads, lo + 2000)
            segs = np.empty((hi - lo, read_len), dtype=np.uint8)
            for k in range(lo, hi):
                d = donors[which[k]]
                st = int(rng.integers(0, len(d) - read_len))
                segs[k - lo] = d[st:st + read_len]
            flat, ends = mutate_reads(rng, segs, err, mix)
            for k in range(lo, hi):
                s = ACGT[flat[ends[k - lo]:ends[k - lo + 1]]]
                if k & 1:
                    s = COMP[s[::-1]]
                f.write(f">read{k}\n".encode() + s.tobytes() + b"\n")
    print(f"[{os.path.basename(outdir)}] data written {time.time() - t0:.0f} s", flush=True)
    subprocess.check_call([exe, "index", "ref.fa"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print(f"[{os.path.basename(outdir)}] index built {time.time() - t0:.0f} s", flush=True)
    with open(os.path.join(tmp, "full.sam"), "w") as f:          # -t N output is identical to -t 1
        subprocess.check_call([exe, "aln", "-t", str(threads), *aln_opts, "ref.fa", "reads.fa"], cwd=tmp, stdout=f,
                              stderr=subprocess.DEVNULL)
    print(f"[{os.path.basename(outdir)}] seeded + aligned by the reference {time.time() - t0:.0f} s", flush=True)
    shutil.rmtree(outdir, ignore_errors=True)
    os.makedirs(outdir)
    for name in ("ref.fa.bwt", "ref.fa.sa", "ref.fa.ann", "ref.fa.amb", "ref.fa.pac", "reads.fa", "reads.fa.seed.gem.map"):
        shutil.move(os.path.join(tmp, name), os.path.join(outdir, name))
    n = 0
    with open(os.path.join(tmp, "full.sam")) as f, open(os.path.join(outdir, "expected.sam"), "w") as g:
        for line in f:
            if not line.startswith("@PG"):
                g.write(line); n += not line.startswith("@")
    with open(os.path.join(outdir, "cmd.txt"), "w") as f:
        f.write(" ".join(aln_opts) + "\n")
    shutil.rmtree(tmp, ignore_errors=True)
    print(f"{outdir}: {n} SAM records, {sum(os.path.getsize(os.path.join(outdir, x)) for x in os.listdir(outdir)) >> 20} MiB, "
          f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    big = os.environ.get("LB2_BIGFIX", os.path.join(ROOT, "bigfix"))      # top level, git- and gpurun-ignored; moved into the snapshot on demand
    for what in sys.argv[1:]:
        if what == "multi":
            d = os.path.join(ROOT, "tests", "golden", "sam_multi")
            make(d, [60_000, 45_000, 80_000, 50_000], 48, 4000, 0.05, seed=21, sv_every=6000, sv_max=1200, transloc=True)
            compress_members(d)
        elif what == "c3full":
            make(os.path.join(big, "sam_c3full"), [100_000_000], 20_000, 10_000, 0.15, seed=31, mix=(1.5, 9, 4.5))
        elif what == "c4full":
            make(os.path.join(big, "sam_c4full"), [100_000_000], 20_000, 20_000, 0.05, seed=41, aln_opts=("-V", "10000"),
                 sv_every=15_000, sv_max=10_000)
        elif what == "c5s":
            make(os.path.join(big, "sam_c5s"), [5_000_000] * 24, 20_000, 10_000, 0.15, seed=51, mix=(1.5, 9, 4.5), transloc=False)
        elif what == "c4multi":
            make(os.path.join(big, "sam_c4multi"), [2_000_000] * 6, 1500, 20_000, 0.05, seed=61, aln_opts=("-V", "10000"),
                 sv_every=15_000, sv_max=10_000, transloc=True)
