"""Builds whole-program SAM fixtures by running the UNMODIFIED reference `lamsa`
(compiled in a scratch copy under /tmp) end to end, GEM seeding included, on
seeded synthetic data (SURVEY.md 8d C1 and reduced C3/C4 shapes).  Test
infrastructure; only possible in the container that has /root/reference.

Each fixture directory holds what `lamsa aln -N` (reuse the GEM map) needs plus
the reference's SAM:  ref.fa.{bwt,sa,ann,amb,pac}  reads.fa  reads.fa.seed.gem.map
expected.sam (without the @PG line)  cmd.txt (extra `lamsa aln` options)

  tests/golden/sam_small/   committed (xz-compressed members)
  oracle/_ref/sam_*/        git-ignored, travels to the GPU box with gpurun

    python oracle/make_sam_fixtures.py
"""
import lzma
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
WORK = "/tmp/lamsa_ref_build"
ACGT = np.array(list("ACGT"))
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def build_reference():
    exe = os.path.join(WORK, "lamsa")
    if not os.path.exists(exe):
        shutil.rmtree(WORK, ignore_errors=True)
        shutil.copytree(REF, WORK)
        subprocess.check_call(["chmod", "-R", "u+w", WORK])
        subprocess.check_call(["make", "-s", "-j8", "-C", WORK,
                               "CFLAGS=-w -O3 -fcommon"])       # -fcommon: SURVEY.md 0.8
    return exe


def mutate(rng, seg, err, mix=(1, 1, 1)):
    s, i, d = (x / sum(mix) for x in mix)
    out = []
    for b in seg:
        x = rng.random()
        if x < err * s:
            out.append((b + rng.integers(1, 4)) % 4)
        elif x < err * (s + d):
            continue
        elif x < err:
            out.append(rng.integers(0, 4)); out.append(b)
        else:
            out.append(b)
    return np.array(out, dtype=np.int64)


def write_fa(path, name, codes):
    s = "".join(ACGT[codes])
    with open(path, "w") as f:
        f.write(f">{name}\n")
        for k in range(0, len(s), 60):
            f.write(s[k:k + 60] + "\n")


def make(outdir, ref_len, n_reads, read_len, err, seed, aln_opts=(), sv=False, mix=(1, 1, 1)):
    exe = build_reference()
    tmp = f"/tmp/lamsa_fixture_{os.path.basename(outdir)}"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, ref_len)
    write_fa(os.path.join(tmp, "ref.fa"), "chr1", ref)
    donor = ref
    if sv:       # donor genome with deletions / insertions / inversions / duplications every ~15 kbp
        parts, pos = [], 0
        while pos < ref_len:
            nxt = min(ref_len, pos + int(rng.integers(10000, 20000)))
            parts.append(ref[pos:nxt])
            if nxt >= ref_len:
                break
            kind, ln = int(rng.integers(0, 4)), int(rng.integers(50, 3000))
            if kind == 0:
                nxt = min(ref_len, nxt + ln)                       # deletion
            elif kind == 1:
                parts.append(rng.integers(0, 4, ln))               # insertion
            elif kind == 2:
                seg = ref[nxt:nxt + ln]; parts.append(3 - seg[::-1]); nxt = min(ref_len, nxt + ln)   # inversion
            else:
                parts.append(ref[max(0, nxt - ln):nxt])            # tandem duplication
            pos = nxt
        donor = np.concatenate(parts)
    with open(os.path.join(tmp, "reads.fa"), "w") as f:
        for r in range(n_reads):
            st = int(rng.integers(0, len(donor) - read_len))
            s = "".join(ACGT[mutate(rng, donor[st:st + read_len], err, mix)])
            if r & 1:
                s = "".join(COMP[c] for c in reversed(s))
            f.write(f">read{r}\n{s}\n")
    subprocess.check_call([exe, "index", "ref.fa"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(os.path.join(tmp, "full.sam"), "w") as f:
        subprocess.check_call([exe, "aln", "-t", "1", *aln_opts, "ref.fa", "reads.fa"], cwd=tmp, stdout=f,
                              stderr=subprocess.DEVNULL)
    os.makedirs(outdir, exist_ok=True)
    for name in ("ref.fa.bwt", "ref.fa.sa", "ref.fa.ann", "ref.fa.amb", "ref.fa.pac", "reads.fa", "reads.fa.seed.gem.map"):
        shutil.copy(os.path.join(tmp, name), os.path.join(outdir, name))
    with open(os.path.join(tmp, "full.sam")) as f, open(os.path.join(outdir, "expected.sam"), "w") as g:
        n = 0
        for line in f:
            if not line.startswith("@PG"):
                g.write(line); n += not line.startswith("@")
    with open(os.path.join(outdir, "cmd.txt"), "w") as f:
        f.write(" ".join(aln_opts) + "\n")
    print(f"{outdir}: {n} SAM records, {sum(os.path.getsize(os.path.join(outdir, x)) for x in os.listdir(outdir)) >> 10} KiB")


def compress_members(d):
    for name in os.listdir(d):
        if name.endswith(".xz") or name == "cmd.txt":
            continue
        p = os.path.join(d, name)
        with open(p, "rb") as f, lzma.open(p + ".xz", "wb", preset=9) as g:
            g.write(f.read())
        os.remove(p)


if __name__ == "__main__":
    small = os.path.join(ROOT, "tests", "golden", "sam_small")
    shutil.rmtree(small, ignore_errors=True)
    make(small, 60_000, 24, 2500, 0.05, seed=5)
    compress_members(small)
    out = os.path.join(ROOT, "oracle", "_ref")
    make(os.path.join(out, "sam_c1"), 1_000_000, 1000, 5000, 0.05, seed=1)                       # BASELINE configs[0]
    make(os.path.join(out, "sam_c3s"), 1_000_000, 100, 10000, 0.15, seed=3, aln_opts=("-T", "pacbio"),
         mix=(1.5, 9, 4.5))                                                                      # configs[2], reduced
    make(os.path.join(out, "sam_c4s"), 1_000_000, 100, 20000, 0.05, seed=4, sv=True)             # configs[3], reduced
    if "--big" in sys.argv:       # throughput fixtures for tools/bench_lamsa.py (minutes of GEM time; xz the members to keep gpurun pushes small)
        make(os.path.join(out, "sam_c1x8"), 4_000_000, 8000, 5000, 0.05, seed=11)
        make(os.path.join(out, "sam_c3m"), 4_000_000, 2000, 10000, 0.15, seed=13, aln_opts=("-T", "pacbio"), mix=(1.5, 9, 4.5))
