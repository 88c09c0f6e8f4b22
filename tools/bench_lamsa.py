#!/usr/bin/env python
"""Whole-program timing of `lamsa aln -N` on a recorded fixture (oracle/make_sam_fixtures.py): the unmodified
reference binary against the batch-producer binary; SAM must be identical.  One JSON line per run with
aligned Mbp/s of the alignment stage and the whole-process wall time.

  python tools/bench_lamsa.py [fixture-dir] [--threads 16] [--in-flight 8192,32768] [--replicate 4]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lamsa_b200 import pipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("fixture", nargs="?", default=os.path.join(ROOT, "oracle", "_ref", "sam_c1"))
    ap.add_argument("--threads", default=str(os.cpu_count()), help="thread counts of the reference runs")
    ap.add_argument("--in-flight", default="8192", help="reads in flight of the producer runs")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--replicate", type=int, default=1, help="concatenate the fixture's reads this many times")
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    work = pipeline.temp_workdir(a.fixture, a.replicate)
    bases = pipeline.read_bases(work)
    exp = list(open(os.path.join(work, "expected.sam")))
    runs = []
    if not a.skip_reference:
        runs += [(pipeline.REFBIN, "reference (CPU ksw.c)", int(t), {}) for t in a.threads.split(",")]
    runs += [(pipeline.PRODUCER, "batch producer (liblamsa_b200, B200)", 1,
              {"LB2_READS_IN_FLIGHT": n, "LB2_DEVICES": str(a.devices), "LB2_FIBER_STATS": "1"}) for n in a.in_flight.split(",")]
    for exe, label, threads, env in runs:
        if not os.path.exists(exe):
            print(json.dumps({"impl": label, "unavailable": exe}))
            continue
        best = None
        for _ in range(a.repeat):
            r = pipeline.run(exe, work, threads, env)
            if best is None or (r["stage_s"] or r["wall_s"]) < (best["stage_s"] or best["wall_s"]):
                best = r
        if env:
            sys.stderr.write("".join(l for l in best["stderr"] if "[lamsa_b200]" in l))
        st = best["stage_s"]
        print(json.dumps({"impl": label, "fixture": os.path.basename(a.fixture), "replicate": a.replicate,
                          "threads": threads, **{k.lower(): v for k, v in env.items() if k != "LB2_FIBER_STATS"},
                          "host_cores": os.cpu_count(), "reads_bases": bases,
                          "wall_s": round(best["wall_s"], 3), "stage_s": round(st, 3) if st else None,
                          "aligned_mbp_per_s": round(bases / st / 1e6, 2) if st else None,
                          "aligned_mbp_per_s_whole_process": round(bases / best["wall_s"] / 1e6, 2),
                          "sam_identical_to_reference": best["sam"] == exp,
                          "records": len([l for l in best["sam"] if not l.startswith("@")])}), flush=True)
    pipeline.cleanup(work)


if __name__ == "__main__":
    main()
