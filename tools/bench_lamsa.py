#!/usr/bin/env python
"""Whole-program timing of `lamsa aln` on a recorded fixture (oracle/make_sam_fixtures.py):
the unmodified reference binary (CPU ksw.c) against the drop-in binary (same program, ksw.c
replaced by liblamsa_b200.so).  Both run with -N (reuse the GEM seed map), so the timed
region is LAMSA's own alignment stage; SAM output must be identical.  Prints one JSON line
per run with aligned Mbp/s = sum(read lengths) / wall seconds.

  python tools/bench_lamsa.py [fixture-dir] [--threads 1,16,128]
"""
import argparse
import json
import lzma
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "lamsa_ref")
DROPIN = os.path.join(ROOT, "oracle", "_ref", os.environ.get("LAMSA_DROPIN", "lamsa_dropin"))


def stage(src, dst):
    os.makedirs(dst)
    for name in os.listdir(src):
        p = os.path.join(src, name)
        if name.endswith(".xz"):
            with lzma.open(p, "rb") as f, open(os.path.join(dst, name[:-3]), "wb") as g:
                g.write(f.read())
        else:
            shutil.copy(p, os.path.join(dst, name))


def run(exe, work, threads, opts):
    out = os.path.join(work, f"out_{os.path.basename(exe)}_{threads}.sam")
    t0 = time.perf_counter()
    with open(out, "w") as f:
        r = subprocess.run([exe, "aln", "-t", str(threads), "-N", *opts, "ref.fa", "reads.fa"], cwd=work, stdout=f,
                           stderr=subprocess.PIPE)
    dt = time.perf_counter() - t0
    if r.returncode:
        raise RuntimeError(r.stderr.decode()[-1000:])
    if os.environ.get("LB2_FIBER_STATS"):
        sys.stderr.write("".join(l + "\n" for l in r.stderr.decode().splitlines() if "[lamsa_b200]" in l))
    return dt, [l for l in open(out) if not l.startswith("@PG")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("fixture", nargs="?", default=os.path.join(ROOT, "oracle", "_ref", "sam_c1"))
    ap.add_argument("--threads", default="1,16,128")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--workers", default="1024,4096", help="worker (fiber) counts of the batch-producer build")
    ap.add_argument("--skip-dropin", action="store_true", help="skip the thread-per-read drop-in build")
    a = ap.parse_args()
    work = os.path.join(tempfile.mkdtemp(prefix="lamsa_bench_"), "w")
    stage(a.fixture, work)
    opts = open(os.path.join(work, "cmd.txt")).read().split()
    bases = sum(len(l.strip()) for l in open(os.path.join(work, "reads.fa")) if not l.startswith(">"))
    exp = list(open(os.path.join(work, "expected.sam")))
    FIBER = os.path.join(ROOT, "oracle", "_ref", "lamsa_dropin_fiber")
    impls = [(REFBIN, "reference (CPU ksw.c)", a.threads)]
    if not a.skip_dropin:
        impls.append((DROPIN, "drop-in, one blocking call per DP task (liblamsa_b200, B200)", a.threads))
    impls.append((FIBER, "batch producer: workers as fibers, DP + chaining batched (liblamsa_b200, B200)", a.workers))
    for exe, label, counts in impls:
        if not os.path.exists(exe):
            print(json.dumps({"impl": label, "unavailable": exe}))
            continue
        for t in [int(x) for x in counts.split(",")]:
            best = None
            for _ in range(a.repeat):
                dt, sam = run(exe, work, t, opts)
                best = dt if best is None else min(best, dt)
            print(json.dumps({"impl": label, "fixture": os.path.basename(a.fixture), "threads": t, "host_cores": os.cpu_count(),
                              "wall_s": round(best, 3), "aligned_mbp_per_s": round(bases / best / 1e6, 2),
                              "sam_identical_to_reference": sam == exp, "records": len([l for l in sam if not l.startswith("@")])}),
                  flush=True)


if __name__ == "__main__":
    main()
