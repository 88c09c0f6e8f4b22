"""Writes tests/golden/sdp_golden.npz: chaining (SDP) inputs with the outputs of the UNMODIFIED
reference frag_line_BCC / frag_line_remain.  Only possible where /root/reference exists.

  groups rec_*  -- call streams recorded from whole `lamsa aln -t 1 -N` runs of the reference on
                   the fixtures of oracle/make_sam_fixtures.py (oracle/_ref/lamsa_rec, built by
                   `make -C oracle sdprec`: the reference program with the two entry points
                   wrapped by the linker);
  groups syn_*  -- synthetic seed-hit sets (tests/_sdp.py:gen_reads) run through
                   oracle/_ref/liblamsa_ref.so (`make -C oracle sdpref`).

    python tests/golden/make_sdp_golden.py
"""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import _sdp  # noqa: E402

out = {}


def add(name, rs, e1, e2):
    out.update(rs.save_dict(name + "/"))
    out[name + "/s1"], out[name + "/o1"], out[name + "/s2"], out[name + "/o2"] = e1[0], e1[1], e2[0], e2[1]
    print(name, len(rs), "reads", len(rs.hits), "hits")


subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sdpref", "sdprec"])
for fx, keep in (("sam_c1", 150), ("sam_c3s", 100), ("sam_c4s", 100)):
    src = os.path.join(ROOT, "oracle", "_ref", fx)
    if not os.path.isdir(src):
        print("skip", fx, "(run oracle/make_sam_fixtures.py first)")
        continue
    work = f"/tmp/sdp_rec_{fx}"
    shutil.rmtree(work, ignore_errors=True)
    shutil.copytree(src, work)
    opts = open(os.path.join(work, "cmd.txt")).read().split()
    env = dict(os.environ, SDP_REC=os.path.join(work, "sdp.rec"))
    with open(os.path.join(work, "out.sam"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "lamsa_rec"), "aln", "-t", "1", "-N", *opts,
                               "ref.fa", "reads.fa"], cwd=work, stdout=f, stderr=subprocess.DEVNULL, env=env)
    rs, e1, e2 = _sdp.parse_recording(env["SDP_REC"])
    idx = np.arange(min(keep, len(rs)))
    sub = rs.subset(idx)
    cut = lambda e: (e[0][:e[1][len(idx)]], e[1][:len(idx) + 1])
    add("rec_" + fx, sub, cut(e1), cut(e2))

for k, (mode, rf, sv, miss) in enumerate((("default", 0.15, 0.3, 0.3), ("pacbio", 0.3, 0.5, 0.2), ("ont2d", 0.1, 0.2, 0.5),
                                           ("default", 0.6, 0.6, 0.05))):
    rs = _sdp.gen_reads(60, seed=100 + k, mode=mode, repeat_frac=rf, sv_rate=sv, miss_frac=miss, read_len=(500, 9000))
    r1, r2, _ = _sdp.ref_run(rs)
    add(f"syn_{k}_{mode}", rs, r1, r2)

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sdp_golden.npz"), **out)
print(os.path.getsize(os.path.join(ROOT, "tests", "golden", "sdp_golden.npz")), "bytes")
