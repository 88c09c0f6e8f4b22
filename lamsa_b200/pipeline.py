"""Whole-program runs of `lamsa aln -N` on a recorded fixture: the reference binary (CPU ksw.c on host threads)
and the batch-producer binary (oracle/_ref/lamsa_b200_aln: the reference program with ksw.c, lamsa_dp_con.c,
lamsa_heap.c replaced by liblamsa_b200.so and the alignment stage by lamsa_b200/host/aln_core.c).  Used by
bench.py's pipeline leg and tools/bench_lamsa.py.

Timing follows SURVEY.md 8(d): aligned Mbp/s = sum of read lengths / wall seconds of the ALIGNMENT STAGE, index
loading and seeding excluded.  The stage is bracketed by the two lines `lamsa aln` itself prints to (unbuffered)
stderr around lamsa_aln_core: "[lamsa_aln] Mapping reads to genome ..." and "[lamsa_aln] Mapping done!"
(reference src/lamsa_aln.c:1268,1273); the whole-process wall time is reported beside it."""
import lzma
import os
import shutil
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "lamsa_ref")
PRODUCER = os.path.join(ROOT, "oracle", "_ref", "lamsa_b200_aln")


def stage(src, dst, replicate=1):
    """Unpack a fixture directory (members may be .xz) into dst.  replicate>1 concatenates the reads, their seed map
    and the expected SAM records `replicate` times (same reference), to make a longer steady-state run."""
    os.makedirs(dst)
    for name in os.listdir(src):
        p = os.path.join(src, name)
        out = os.path.join(dst, name[:-3] if name.endswith(".xz") else name)
        data = lzma.open(p, "rb").read() if name.endswith(".xz") else open(p, "rb").read()
        base = os.path.basename(out)
        if replicate > 1 and base in ("reads.fa", "reads.fa.seed.gem.map"):
            data = data * replicate
        elif replicate > 1 and base == "expected.sam":
            lines = data.splitlines(keepends=True)
            head = [l for l in lines if l.startswith(b"@")]
            body = [l for l in lines if not l.startswith(b"@")]
            data = b"".join(head + body * replicate)
        with open(out, "wb") as g:
            g.write(data)
    return dst


def read_bases(work):
    return sum(len(l.strip()) for l in open(os.path.join(work, "reads.fa")) if not l.startswith(">"))


def run(exe, work, threads, env=None, timeout=3600):
    """One `lamsa aln -t threads -N` run -> dict(wall_s, stage_s, sam lines without @PG, stderr tail)."""
    opts = open(os.path.join(work, "cmd.txt")).read().split()
    out = os.path.join(work, f"out_{os.path.basename(exe)}.sam")
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    with open(out, "w") as f:
        p = subprocess.Popen([exe, "aln", "-t", str(threads), "-N", *opts, "ref.fa", "reads.fa"], cwd=work, stdout=f,
                             stderr=subprocess.PIPE, env=e)
        t_begin = t_end = None
        tail = []
        for raw in p.stderr:                       # the program's own progress lines, time-stamped as they arrive
            now = time.perf_counter()
            line = raw.decode(errors="replace")
            if "Mapping reads to genome" in line:
                t_begin = now
            elif "Mapping done" in line:
                t_end = now
            tail.append(line)
            tail = tail[-40:]
        rc = p.wait(timeout=timeout)
    wall = time.perf_counter() - t0
    if rc:
        raise RuntimeError(f"{exe} exited {rc}: {''.join(tail)[-2000:]}")
    sam = [l for l in open(out) if not l.startswith("@PG")]
    return {"wall_s": wall, "stage_s": (t_end - t_begin) if t_begin and t_end else None, "sam": sam, "stderr": tail}


def temp_workdir(fixture, replicate=1):
    return stage(fixture, os.path.join(tempfile.mkdtemp(prefix="lamsa_pipe_"), "w"), replicate)


def cleanup(work):
    shutil.rmtree(os.path.dirname(work), ignore_errors=True)
