#!/usr/bin/env python
"""Run `lamsa aln -N` of the reference binary and of the batch producer on a packed big fixture
(tools/pack_bigfix.py) and compare each SAM's sha1 with the reference's recorded one.  One JSON line per run.
   python tools/run_bigfix.py <packed-dir> [--in-flight N] [--devices N] [--skip-reference]"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lamsa_b200 import pipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("fixture")
ap.add_argument("--in-flight", default="8192")
ap.add_argument("--devices", type=int, default=1)
ap.add_argument("--skip-reference", action="store_true")
a = ap.parse_args()
work = os.path.join(tempfile.mkdtemp(prefix="lamsa_big_", dir=os.environ.get("LB2_TMP", None)), "w")
os.makedirs(work)
t0 = time.time()
procs = []
for name in os.listdir(a.fixture):
    p = os.path.join(a.fixture, name)
    if name.endswith(".xz"):
        procs.append(subprocess.Popen(["xz", "-T0", "-d", "-c", p], stdout=open(os.path.join(work, name[:-3]), "wb")))
    else:
        subprocess.check_call(["cp", p, os.path.join(work, name)])
for p in procs:
    p.wait()
exp_sha, exp_n = open(os.path.join(work, "expected.sha1")).read().split()
open(os.path.join(work, "expected.sam"), "w").close()
bases = pipeline.read_bases(work)
sys.stderr.write(f"staged in {time.time() - t0:.1f} s, {bases} bases\n")
runs = [] if a.skip_reference else [(pipeline.REFBIN, "reference (CPU ksw.c)", os.cpu_count(), {})]
runs += [(pipeline.PRODUCER, "batch producer (liblamsa_b200, B200)", 1,
          {"LB2_READS_IN_FLIGHT": n, "LB2_DEVICES": str(a.devices), "LB2_FIBER_STATS": "1"}) for n in a.in_flight.split(",")]
for exe, label, threads, env in runs:
    r = pipeline.run(exe, work, threads, env)
    h, n = hashlib.sha1(), 0
    for l in r["sam"]:
        h.update(l.encode()); n += not l.startswith("@")
    if env:
        sys.stderr.write("".join(l for l in r["stderr"] if "[lamsa_b200]" in l))
    print(json.dumps({"impl": label, "fixture": os.path.basename(os.path.normpath(a.fixture)), "threads": threads,
                      **{k.lower(): v for k, v in env.items() if k != "LB2_FIBER_STATS"}, "host_cores": os.cpu_count(),
                      "reads_bases": bases, "wall_s": round(r["wall_s"], 3), "stage_s": round(r["stage_s"], 3),
                      "aligned_mbp_per_s": round(bases / r["stage_s"] / 1e6, 2), "records": n, "expected_records": int(exp_n),
                      "sam_sha1_identical_to_reference": h.hexdigest() == exp_sha}), flush=True)
pipeline.cleanup(work)
