"""GPU suite: whole-program drop-in check.  oracle/_ref/lamsa_dropin is the
reference's own `lamsa` with ONLY ksw.c replaced by liblamsa_b200.so (built by
`make -C oracle dropin` where the reference tree exists).  It is run with `-N`
(reuse the recorded GEM seed map) on fixtures produced by the unmodified
reference (oracle/make_sam_fixtures.py) and its SAM must be identical, record
for record, to the reference's (only the @PG command line is excluded)."""
import lzma
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# lamsa_dropin: only ksw.c replaced; lamsa_dropin_sdp: ksw.c, lamsa_dp_con.c and lamsa_heap.c replaced
# (banded DP and sparse-DP chaining both on the GPU; `make -C oracle dropin_sdp`)
EXES = {"dp": os.path.join(ROOT, "oracle", "_ref", "lamsa_dropin"),
        "dp+sdp": os.path.join(ROOT, "oracle", "_ref", "lamsa_dropin_sdp"),
        # + the alignment stage replaced by this repo's read pipeline (lamsa_b200/host/aln_core.c): reads as worker
        # fibers, every DP / chaining call served in batches gathered over all reads (`make -C oracle producer`)
        "producer": os.path.join(ROOT, "oracle", "_ref", "lamsa_b200_aln"),
        # + lamsa_res_aux (NM / AS of every record) walking the CIGARs on the GPU over the resident reference
        # (lamsa_b200/host/res_aux.c, `make -C oracle producer_aux`): the NM:i / AS:i tags must not move
        "producer+aux": os.path.join(ROOT, "oracle", "_ref", "lamsa_b200_aln_aux"),
        # + init_hash / hash_split_map of the local split mapping: k-mer index of the window, look-up and chaining of the
        # hits on the GPU (hash_line.cuh, lamsa_b200/host/split_map.c, `make -C oracle producer_hash`)
        "producer+hash": os.path.join(ROOT, "oracle", "_ref", "lamsa_b200_aln_hash")}
FIXTURES = [
    ("small", os.path.join(ROOT, "tests", "golden", "sam_small")),
    # 4 contigs; donor with deletions / insertions / inversions / duplications and translocations BETWEEN contigs
    # (oracle/make_big_fixtures.py multi): E_CHR_DIF edges of the chaining and records on several contigs per read
    ("multi_contig_sv", os.path.join(ROOT, "tests", "golden", "sam_multi")),
    ("c1", os.path.join(ROOT, "oracle", "_ref", "sam_c1")),
    ("c3_reduced_pacbio", os.path.join(ROOT, "oracle", "_ref", "sam_c3s")),
    ("c4_reduced_sv", os.path.join(ROOT, "oracle", "_ref", "sam_c4s")),
]


def stage(src, dst):
    os.makedirs(dst)
    for name in os.listdir(src):
        p = os.path.join(src, name)
        if name.endswith(".xz"):
            with lzma.open(p, "rb") as f, open(os.path.join(dst, name[:-3]), "wb") as g:
                g.write(f.read())
        else:
            shutil.copy(p, os.path.join(dst, name))


@pytest.mark.parametrize("name,src", FIXTURES, ids=[f[0] for f in FIXTURES])
@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("link", ["dp", "dp+sdp", "producer", "producer+aux", "producer+hash"])
def test_dropin_sam_identical_to_reference(tmp_path, name, src, threads, link):
    EXE = EXES[link]
    if not os.path.exists(EXE):
        pytest.skip(f"{EXE} not built (needs the reference tree at build time)")
    if not os.path.isdir(src):
        pytest.skip(f"fixture {src} not present")
    env = dict(os.environ)
    if link.startswith("producer"):
        env["LB2_FIBER_STATS"] = "1"
        # reads in flight: the default (thousands), or fewer workers than reads with an odd count
        if threads == 1:
            env["LB2_READS_IN_FLIGHT"] = "37"
    elif threads != 1 and name not in ("small", "c1", "multi_contig_sv"):
        pytest.skip("multi-thread run only on two fixtures")
    work = str(tmp_path / name)
    stage(src, work)
    opts = open(os.path.join(work, "cmd.txt")).read().split()
    with open(os.path.join(work, "out.sam"), "w") as f:
        r = subprocess.run([EXE, "aln", "-t", str(threads), "-N", *opts, "ref.fa", "reads.fa"], cwd=work, stdout=f,
                           stderr=subprocess.PIPE, timeout=1200, env=env)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = [l for l in open(os.path.join(work, "out.sam")) if not l.startswith("@PG")]
    exp = list(open(os.path.join(work, "expected.sam")))
    assert len(got) == len(exp), (len(got), len(exp))
    diff = [(i, a, b) for i, (a, b) in enumerate(zip(got, exp)) if a != b]
    assert not diff, f"{len(diff)} differing SAM lines; first: {diff[0][1][:300]!r} vs {diff[0][2][:300]!r}"
    if link == "producer+hash" and name in ("multi_contig_sv", "c4_reduced_sv"):
        assert b"split-mapping lines" in r.stderr or "LB2_FIBER_STATS" not in env      # the SV fixtures do reach the GPU path
