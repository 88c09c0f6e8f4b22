"""CPU suite: the read pipeline of lamsa_b200/host/aln_core.c (this repo's lamsa_aln_core: reads taken one at a
time by worker fibers, no chunk barrier, SAM records formatted by the workers and written in input order) over the
reference's own CPU ksw.c / lamsa_dp_con.c (oracle/_ref/lamsa_pipeline_cpu, `make -C oracle pipeline_cpu`).  No
request reaches the GPU library in this build -- it only provides the fibers -- so what is checked here is the host
logic: the link-time replacement of the reference's lamsa_aln_core, the reader, the workers and the in-order writer.
SAM must be identical to the unmodified reference's.  Only where the reference tree was present at build time."""
import os
import subprocess

import pytest

from test_gpu_dropin_sam import FIXTURES, ROOT, stage

EXE = os.path.join(ROOT, "oracle", "_ref", "lamsa_pipeline_cpu")
pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/lamsa_pipeline_cpu not built (needs /root/reference)")


@pytest.mark.parametrize("name,src", FIXTURES, ids=[f[0] for f in FIXTURES])
@pytest.mark.parametrize("threads,in_flight", [(1, 1), (1, 7), (8, 61)])
def test_pipeline_sam_identical_to_reference(tmp_path, name, src, threads, in_flight):
    if not os.path.isdir(src):
        pytest.skip(f"fixture {src} not present")
    if name not in ("small", "c4_reduced_sv") and in_flight == 1:
        pytest.skip("single-worker run only on two fixtures")
    work = str(tmp_path / name)
    stage(src, work)
    opts = open(os.path.join(work, "cmd.txt")).read().split()
    env = dict(os.environ, LB2_HOST_THREADS=str(threads), LB2_READS_IN_FLIGHT=str(in_flight))
    with open(os.path.join(work, "out.sam"), "w") as f:
        r = subprocess.run([EXE, "aln", "-t", "1", "-N", *opts, "ref.fa", "reads.fa"], cwd=work, stdout=f,
                           stderr=subprocess.PIPE, timeout=1200, env=env)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = [l for l in open(os.path.join(work, "out.sam")) if not l.startswith("@PG")]
    exp = list(open(os.path.join(work, "expected.sam")))
    assert got == exp
