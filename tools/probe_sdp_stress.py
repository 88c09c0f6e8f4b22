"""Manual GPU triage for the chaining kernel: heavy multi-hit reads (up to per_aln_m hits per seed,
most seeds repetitive) against the oracle; meant to be run under compute-sanitizer as well:
    compute-sanitizer --tool memcheck python tools/probe_sdp_stress.py 40
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import _sdp
import lamsa_b200
from lamsa_b200.sdp import SdpBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = lamsa_b200.Context(0)
bad = []
for k, (mode, rf, mh) in enumerate((("default", 0.9, 200), ("pacbio", 0.7, 120), ("ont2d", 0.5, 200), ("default", 0.3, 30))):
    rs = _sdp.gen_reads(n, seed=50 + k, mode=mode, repeat_frac=rf, sv_rate=0.5, miss_frac=0.1, max_hits=mh, read_len=(2000, 9000))
    t0 = time.time(); o1, o2, op = _sdp.oracle_run(rs); t_o = time.time() - t0
    b = SdpBatch(ctx, rs.para, rs.reads, rs.seed_id, rs.map_n, rs.hits)
    t0 = time.time(); g1 = b.run_bcc(); k1 = b.kernel_ms; p1 = b.stats()["pairs"]; g2 = b.run_remain(rs.reads, rs.regs); t_g = time.time() - t0
    p2 = b.stats()["pairs"]; b.close()
    d = _sdp.diff_streams(g1, o1, f"{mode} stage1") + _sdp.diff_streams(g2, o2, f"{mode} stage2")
    print(mode, "reads", len(rs), "hits", len(rs.hits), "max hits/read", int(max(rs.map_n.sum() for _ in [0])) if False else "",
          "pairs", int(op.sum()), (p1, p2), "oracle %.2fs gpu %.2fs" % (t_o, t_g), "mismatches", len(d))
    bad += d
ctx.close()
if bad:
    print("\n".join(bad[:10])); sys.exit(1)
print("chaining stress ok")
