"""Manual GPU triage: kernel time of ONE long task (fill + traceback), the quantity that bounds a side batch of the producer."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lamsa_b200
from lamsa_b200 import workload
ctx = lamsa_b200.Context(0)
for rows, w, kind in ((4700, 10, 1), (4700, 50, 1), (4700, 10, 0), (19000, 10, 1), (2000, 200, 1)):
    tasks, keep = workload.gen_microbench(2, seed=11, qmin=rows, qmax=rows, wmin=w, wmax=w, max_err=0.05, max_dl=5)
    tasks = tasks[tasks["kind"] == kind][:1].copy()
    tasks["h0"] = 100 if kind else 0
    b = lamsa_b200.Batch(ctx, tasks, keep); b.upload(); b.compute(); b.compute()
    ms = b.compute(); st = b.stats(); res, cig = b.download()
    print(json.dumps({"rows": rows, "w": w, "kind": kind, "rows_done": int(res["tle"][0]), "cells": int(res["cells"][0]), "n_cigar": int(res["n_cigar"][0]),
                      "fill_ms": round(st["fill_ms"], 3), "trace_ms": round(st["trace_ms"], 3),
                      "fill_us_per_row": round(1e3 * st["fill_ms"] / max(1, int(res["tle"][0])), 3),
                      "trace_us_per_row": round(1e3 * st["trace_ms"] / max(1, int(res["tle"][0])), 3)}), flush=True)
    b.close()
