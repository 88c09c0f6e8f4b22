#!/usr/bin/env python
"""Make a big SAM fixture (oracle/make_big_fixtures.py) small enough to travel with a gpurun snapshot: members
xz-compressed in parallel, and the reference's expected.sam replaced by its record count and sha1 (over every line
but @PG) -- tools/run_bigfix.py compares digests.   python tools/pack_bigfix.py <src-dir> <dst-dir>"""
import hashlib
import os
import shutil
import subprocess
import sys

src, dst = sys.argv[1], sys.argv[2]
shutil.rmtree(dst, ignore_errors=True)
os.makedirs(dst)
h, n = hashlib.sha1(), 0
with open(os.path.join(src, "expected.sam"), "rb") as f:
    for line in f:
        if not line.startswith(b"@PG"):
            h.update(line)
            n += not line.startswith(b"@")
with open(os.path.join(dst, "expected.sha1"), "w") as f:
    f.write(f"{h.hexdigest()} {n}\n")
procs = []
for name in os.listdir(src):
    p = os.path.join(src, name)
    if name == "expected.sam":
        continue
    if os.path.getsize(p) < 1 << 20:
        shutil.copy(p, os.path.join(dst, name))
    else:
        out = open(os.path.join(dst, name + ".xz"), "wb")
        procs.append(subprocess.Popen(["xz", "-T0", "-2", "-c", p], stdout=out))
for p in procs:
    p.wait()
print(dst, sum(os.path.getsize(os.path.join(dst, x)) for x in os.listdir(dst)) >> 20, "MiB")
