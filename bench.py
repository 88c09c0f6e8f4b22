#!/usr/bin/env python
"""bench.py -- the two throughput figures of BASELINE.json's metric, on B200:

  1. banded-DP GCUPS on BASELINE configs[1] (the ksw_global2 / ksw_extend_core microbenchmark: 1 M synthetic
     DP tasks, 50-1000 bp, band 10..200, half global / half extension, CIGARs produced) -- the JSON line's
     `value`, `e2e`, `roofline`;
  2. aligned Mbp/s of the whole alignment stage of `lamsa aln` (sum of read lengths / wall seconds
     of the stage, seeding and index loading excluded) on a C3-shaped fixture -- the line's `pipeline` object and
     `aligned_mbp_per_s`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--tasks T] [--impl reference]

A "step" = one pass of the hot path (fill + traceback kernels) over the whole task batch, inputs resident in
HBM.  `value` = DP cells the reference would evaluate (counted by the kernels, identical to the oracle's count) /
device time, summed over ranks; `e2e` = the same through lb2_dp_run_pool with HOST task records and the sequences in one page-locked host pool
(classification + H2D of the pool as it lies + device re-layout + kernels + D2H of scores/CIGARs), timed over the same number of steps.  N>1: one process per GPU (torchrun), tasks sharded
by rank (independent batches, no collective on the data path), weak scaling; the pipeline leg then runs ONE
process that drives all N GPUs (reads are spread over the GPUs by the batch producer).
`--impl reference` times the reference's own CPU code on all host cores: ksw.c (oracle/_ref/libksw_ref.so, or the
oracle port when that was not built) on a bounded sample of the DP workload, and the unmodified `lamsa aln`
(oracle/_ref/lamsa_ref) on the same pipeline fixture.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

OPS_EXTEND, OPS_GLOBAL = 22, 16        # algorithmic integer ops per cell
METRIC = "banded_dp_gcups"
DTYPE = "s16x2 (packed int16 DPX lanes where the value range is provably safe, int32 lanes otherwise; bit-exact with the reference's int32)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p = gpu, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the upper half of samples = clocks under load (idle samples sit at the floor)
        hi = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(tasks, threads):
    """Time the reference ksw.c (or the oracle port) on host threads -> (gcups, kind, seconds, cells)."""
    import _oracle
    ores, _, _ = _oracle.oracle_run(tasks[:1], 1)      # builds/loads the oracle
    if _oracle.have_ref():
        kind = "reference"
        res, cig, secs = _oracle.ref_run(tasks, threads)
        cells = int(_oracle.oracle_run(tasks, threads)[0]["cells"].sum())   # the reference .so has no cell counter
    else:
        kind = "port"
        res, cig, secs = _oracle.oracle_run(tasks, threads)
        cells = int(res["cells"].sum())
    return cells / secs / 1e9, kind, secs, cells


# ------------------------------------------------------------------- pipeline leg
def pipeline_fixture():
    """(dir, replicate, description) of the C3-shaped fixture the pipeline leg runs, or None.  The fixtures are outputs
    of the unmodified reference (oracle/make_sam_fixtures.py); the larger ones live under oracle/_ref/ (not in git,
    shipped with the snapshot), the committed small one is the fall-back."""
    cands = [(os.path.join(ROOT, "oracle", "_ref", "sam_c3m"), 10,
              "C3 shape at BASELINE configs[2]'s read count: 4 Mbp random reference, 2 000 x 10 kbp reads at 15 % error (1.5/9/4.5 sub/ins/del), `-T pacbio`, reads taken 10 times = 20 000 reads, 209 Mbp"),
             (os.path.join(ROOT, "oracle", "_ref", "sam_c3s"), 16,
              "C3 shape, reduced: 1 Mbp reference, 100 x 10 kbp reads at 15 % error, `-T pacbio`, reads taken 16 times"),
             (os.path.join(ROOT, "tests", "golden", "sam_small"), 64,
              "committed fall-back: 60 kbp reference, 24 x 2.5 kbp reads at 5 % error, reads taken 64 times")]
    for d, rep, desc in cands:
        if os.path.isdir(d):
            return d, env_int("LB2_PIPE_REPLICATE", rep), desc
    return None


def pipeline_replicate(base, n_gpus):
    """Weak scaling like the DP leg: the reads are taken `base` times per GPU."""
    return base * max(1, n_gpus)


def pipeline_leg(impl, n_gpus, repeat=2):
    """Whole `lamsa aln -N` on the pipeline fixture -> dict for the JSON line (never raises)."""
    try:
        from lamsa_b200 import pipeline
        fx = pipeline_fixture()
        exe = pipeline.REFBIN if impl == "reference" else pipeline.PRODUCER
        if fx is None or not os.path.exists(exe):
            return {"unavailable": f"{'fixture' if fx is None else exe} not present (built where the reference tree is mounted)"}
        rep = pipeline_replicate(fx[1], n_gpus)
        work = pipeline.temp_workdir(fx[0], rep)
        bases = pipeline.read_bases(work)
        exp = list(open(os.path.join(work, "expected.sam")))
        cores = os.cpu_count() or 1
        in_flight = env_int("LB2_READS_IN_FLIGHT", min(8192 * n_gpus, 24576))
        env = {} if impl == "reference" else {"LB2_DEVICES": str(n_gpus), "LB2_READS_IN_FLIGHT": str(in_flight), "LB2_FIBER_STATS": "1",
                                              "LB2_READ_TRACE": os.path.join(work, "read_trace.txt")}
        best, steady = None, None
        for _ in range(repeat):
            r = pipeline.run(exe, work, cores if impl == "reference" else 1, env)
            if best is None or r["stage_s"] < best["stage_s"]:
                best = r
                steady = None
                if env:       # steady state: reads completed between 10 % and 90 % of the run, from the per-read trace
                    try:
                        t_end = np.sort(np.loadtxt(os.path.join(work, "read_trace.txt"))[:, 2])
                        lo, hi = int(0.1 * len(t_end)), int(0.9 * len(t_end))
                        if hi > lo and t_end[hi] > t_end[lo]:
                            steady = (hi - lo) / len(t_end) * bases / (t_end[hi] - t_end[lo]) / 1e6
                    except Exception:
                        steady = None
        out = {"metric": "aligned_mbp_per_s", "value": bases / best["stage_s"] / 1e6, "unit": "Mbp/s",
               "definition": "sum of read lengths / wall seconds of the alignment stage (lamsa_aln_core), seeding and index loading excluded (SURVEY.md 8d); "
                             "the stage is bracketed by the program's own stderr lines 'Mapping reads to genome' / 'Mapping done'",
               "stage_s": best["stage_s"], "whole_process_s": best["wall_s"], "whole_process_mbp_per_s": bases / best["wall_s"] / 1e6,
               "fixture": fx[2] + (f"; x{n_gpus} for {n_gpus} GPUs (weak scaling: the same reads per GPU)" if n_gpus > 1 else ""), "reads_bases": bases, "sam_records": len([l for l in best["sam"] if not l.startswith("@")]),
               "sam_identical_to_reference": best["sam"] == exp, "runs": repeat, "host_cores": cores}
        if impl == "reference":
            out["program"] = f"unmodified reference lamsa aln -t {cores} -N (oracle/_ref/lamsa_ref)"
        else:
            out["program"] = ("reference lamsa aln -N with ksw.c / lamsa_dp_con.c / lamsa_heap.c replaced by liblamsa_b200.so and the alignment "
                              "stage by lamsa_b200/host/aln_core.c (oracle/_ref/lamsa_b200_aln)")
            out["n_gpus"] = n_gpus
            out["reads_in_flight"] = in_flight
            out["steady_state_mbp_per_s"] = steady
            out["stats"] = [l.strip() for l in best["stderr"] if "[lamsa_b200]" in l]
        pipeline.cleanup(work)
        return out
    except Exception as e:
        return {"error": repr(e)[:500]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tasks", type=int, default=env_int("LB2_BENCH_TASKS", 1_000_000), help="DP tasks per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (default: --steps)")
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--no-pipeline", action="store_true", help="skip the whole-program leg")
    a = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world > 1 and "LB2_HOST_THREADS" not in os.environ:      # ranks share the host cores for packing
        os.environ["LB2_HOST_THREADS"] = str(max(8, (os.cpu_count() or 1) // world))
    from lamsa_b200 import workload
    workload_name = (f"ksw microbenchmark C2: {a.tasks} tasks/GPU, qlen U[50,1000], w U[10,200], "
                     "half ksw_global2 / half ksw_extend_core, CIGAR on")
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        n = min(a.tasks, a.cpu_sample)
        tasks, keep = workload.gen_microbench(n, seed=a.seed)
        vals, secs_all = [], []
        kind = "port"
        for s in range(a.warmup + a.steps):
            g, kind, secs, cells = cpu_reference(tasks, threads)
            if s >= a.warmup:
                vals.append(g); secs_all.append(secs)
        v = float(np.mean(vals))
        pipe = None if a.no_pipeline else pipeline_leg("reference", a.gpus)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(secs_all)),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": {"workload": workload_name, "sample_tasks": n},
                "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": threads, "kind": kind,
                                 "sample": f"first {n} tasks of the workload, pthread pool over all host cores"},
                "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if pipe is not None:
            line["pipeline"] = pipe
            line["aligned_mbp_per_s"] = pipe.get("value")
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import lamsa_b200
    dist = None
    if world > 1:
        # NCCL only carries the timing barrier and the reduction of the reported numbers (no data-path collective)
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU group for the wait during the whole-program leg: an NCCL barrier would keep the idle ranks spinning on
        # the cores the leg's host threads need
        cpu_group = dist.new_group(backend="gloo")
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    tasks, keep = workload.gen_microbench(a.tasks, seed=a.seed + 1000 * rank)
    ctx = lamsa_b200.Context(dev)
    peak = ctx.int_peak()
    batch = lamsa_b200.Batch(ctx, tasks, keep)
    batch.upload()
    for _ in range(max(a.warmup, 3)):
        batch.compute()
    sampler = ClockSampler(dev)
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    dev_ms, fill_ms, trace_ms, launches = 0.0, 0.0, 0.0, 0
    for _ in range(a.steps):
        dev_ms += batch.compute()                    # CUDA events on the library's stream
        st = batch.stats()
        fill_ms += st["fill_ms"]; trace_ms += st["trace_ms"]; launches += st["launches"]
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    res, cig = batch.download()
    cells = int(res["cells"].sum())
    ext = tasks["kind"] == 1
    cells_ext = int(res["cells"][ext].sum())
    ops_per_step = cells_ext * OPS_EXTEND + (cells - cells_ext) * OPS_GLOBAL
    # algorithmic HBM bytes of one step: sequences in, direction nibbles (+ row bands) out and back in
    # along the traceback path, CIGAR words out
    seq_bytes = int(tasks["qlen"].sum() + tasks["tlen"].sum())
    dir_bytes = cells // 2
    cigar_bytes = int(res["n_cigar"].sum()) * 4
    # one more, untimed pass with the kernel classes launched ONE AFTER THE OTHER, each with its own CUDA events:
    # per kernel its cells, its time alone and its fraction of the integer roof (in a step the classes overlap)
    classes = []
    if rank == 0:
        batch.set_class_timing(True)
        batch.compute()
        batch.download()
        peak_ops = peak["gops_s16x2"] * 1e9
        for c in batch.class_stats():
            opc = OPS_EXTEND if c["kind"] == "extend" else OPS_GLOBAL
            c["gcups"] = c["cells"] / (c["ms"] * 1e-3) / 1e9 if c["ms"] > 0 else None
            c["frac_of_int_roof"] = c["cells"] * opc / (c["ms"] * 1e-3) / peak_ops if c["ms"] > 0 else None
            c["dir_bytes_written"] = c["cells"] // 2
            classes.append(c)
        classes.sort(key=lambda c: -c["ms"])
    batch.close()

    # ---- end to end through the public one-shot call (lb2_dp_run): host task records in, host
    # results + CIGAR words out; packing, H2D and D2H are inside the timed region
    e2e_steps = a.e2e_steps or a.steps
    e2e_secs, h2d, d2h = [], 0, 0
    # the inputs of a step wait in page-locked host memory (one pool holding every sequence, the task records
    # pointing into it): lb2_dp_run_pool uploads the pool as it lies, chunk by chunk, and lays it out on the device
    ptasks, ppool = workload.pool_tasks(tasks, keep, lamsa_b200.pinned_pool)
    for s in range(1 + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        r2, c2 = ctx.run_pool(ptasks, ppool)
        barrier()
        if s >= 1:
            e2e_secs.append(time.perf_counter() - t0)
        st2 = ctx.last_run_stats()
        h2d, d2h = st2["h2d_bytes"], st2["d2h_bytes"]
        del r2, c2
    e2e_s = float(np.mean(e2e_secs))

    # ---- reduce over ranks: max time, summed cells
    from lamsa_b200 import sharding
    (tmax, e2e_max, fill_max), (cells_all, ops_all) = sharding.reduce_metrics(
        dist, "cuda", [dev_ms, e2e_s, fill_ms], [cells, ops_per_step])
    cells_all, ops_all = int(cells_all), int(ops_all)
    ms_step = tmax / a.steps
    value = cells_all / (ms_step * 1e-3) / 1e9

    cpu = None
    if rank == 0 and world == 1:
        n = min(a.tasks, a.cpu_sample)
        g, kind, secs, ccells = cpu_reference(tasks[:n], threads)
        cpu = {"value": g, "unit": "GCUPS", "cores": threads, "kind": kind,
               "sample": f"first {n} tasks of the same workload ({ccells} cells, {secs:.2f} s wall), pthread pool"}

    # ---- secondary leg: the sparse-DP chaining of the same hot path, bounded to a few seconds
    sdp = None
    if rank == 0 and world == 1:
        try:
            import _sdp
            from lamsa_b200.sdp import SdpBatch
            rs = _sdp.gen_reads(8000, seed=7, mode="pacbio", repeat_frac=0.2, sv_rate=0.3, miss_frac=0.3, read_len=(8000, 12000))
            best_ms, pairs = None, 0
            for _ in range(3):
                t0 = time.perf_counter()
                sb = SdpBatch(ctx, rs.para, rs.reads, rs.seed_id, rs.map_n, rs.hits)
                sb.run_bcc(); k = sb.kernel_ms; p1 = sb.stats()["pairs"]
                sb.run_remain(rs.reads, rs.regs); k += sb.kernel_ms; p2 = sb.stats()["pairs"]
                e2e_sdp = time.perf_counter() - t0          # pack + H2D + both stages + D2H of the skeleton streams
                sb.close()
                if best_ms is None or k < best_ms:
                    best_ms, pairs, best_e2e = k, p1 + p2, e2e_sdp
            sub = rs.subset(np.arange(300))
            use_ref = _sdp.have_ref()
            t0 = time.perf_counter()
            _, _, op = (_sdp.ref_run(sub) if use_ref else _sdp.oracle_run(sub))
            t_cpu = time.perf_counter() - t0
            if use_ref:      # the reference library has no pair counter: take the count from the port (same pairs by construction)
                op = _sdp.oracle_run(sub)[2]
            sdp = {"workload": f"{len(rs)} reads x 10 kbp, -T pacbio seeds, {len(rs.hits)} hits, both chaining stages",
                   "gpairs_per_s_kernel": pairs / (best_ms * 1e-3) / 1e9, "reads_per_s_kernel": len(rs) / (best_ms * 1e-3),
                   "reads_per_s_e2e": len(rs) / best_e2e, "unit": "predecessor pairs classified (src/lamsa_dp_con.c:713-751)",
                   "cpu_baseline": {"kind": "reference" if use_ref else "port", "cores": 1, "gpairs_per_s": float(np.sum(op)) / t_cpu / 1e9,
                                    "reads_per_s": len(sub) / t_cpu,
                                    "sample": "first 300 reads, " + ("oracle/_ref/liblamsa_ref.so (unmodified lamsa_dp_con.c)" if use_ref else "oracle/sdp_oracle.c")}}
        except Exception as e:      # the leg is informative; the headline metric does not depend on it
            sdp = {"error": repr(e)}

    ctx.close()
    barrier()
    pipe = None
    if rank == 0 and not a.no_pipeline:
        torch.cuda.empty_cache()
        pipe = pipeline_leg("b200", world)
    if dist is not None:
        dist.barrier(group=cpu_group)

    traffic, traffic_note = None, "no ncu capture committed"
    try:        # dram bytes of the dominant kernel from the committed ncu capture, scaled to this task count
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        tj = json.load(open(tp if os.path.exists(tp) else os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * a.tasks / tj["tasks"]
        traffic_note = f"{tj['dominant_kernel']}: dram read+write per launch from {tj['source']}, scaled x{a.tasks / tj['tasks']:g}"
        if "algorithmic_bytes" in tj:
            traffic_note += (f"; algorithmic bytes of that launch {tj['algorithmic_bytes']:.3e} (ratio {(tj['dram_bytes_read'] + tj['dram_bytes_write']) / tj['algorithmic_bytes']:.2f}), "
                             f"{tj.get('warp_instructions_per_cell')} warp instructions per cell")
    except Exception:
        pass
    if rank == 0:
        peak_tops = peak["gops_s16x2"] / 1e3
        fill_step_ms = fill_max / a.steps
        achieved = (ops_all / world) / (fill_step_ms * 1e-3) / 1e12
        dom = classes[0] if classes else None
        line = {
            "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": workload_name, "tasks_per_gpu": a.tasks, "cells_per_gpu": cells,
                       "l2": "inputs (sequence pool + direction scratch) far larger than the 126 MB L2",
                       "parallelism": f"task-sharded x{world}, no collective"},
            "wall_ms_per_step": 1e3 * t_wall / a.steps,
            "e2e": {"value": cells_all / e2e_max / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "seconds_per_step": e2e_max, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "int_alu", "achieved": achieved, "peak": peak_tops, "unit": "Tops/s (int16 lane-ops)",
                         "frac": achieved / peak_tops,
                         "peak_source": "measured live: VIADDMNMX.S16x2 issue rate (lb2_int_peak); "
                                        "MEASURED_PEAKS.json has no integer figure",
                         "scope": "all fill kernels of a step (they run concurrently on side streams): ops of every cell / CUDA-event time from the first fill launch to the last one's end",
                         "kernel": dom["kernel"] if dom else None,
                         "kernel_alone": dom,
                         "kernel_ms_per_step": fill_step_ms,
                         "trace_ms_per_step": trace_ms / a.steps,
                         "ops_per_cell": {"extend": OPS_EXTEND, "global": OPS_GLOBAL},
                         "classes_alone": classes,
                         "hbm": {"algorithmic_bytes": seq_bytes + 2 * dir_bytes + cigar_bytes,
                                 "achieved_gbs": (seq_bytes + 2 * dir_bytes + cigar_bytes) / (ms_step * 1e-3) / 1e9},
                         "traffic": traffic, "traffic_note": traffic_note},
            "int_peak": peak,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if sdp:
            line["sdp"] = sdp
        if pipe is not None:
            line["pipeline"] = pipe
            line["aligned_mbp_per_s"] = pipe.get("value")
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
