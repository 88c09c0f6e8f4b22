#!/bin/bash
# Producer knob sweep on one fixture: prints stage seconds and the library's own statistics per setting.
# usage: tools/producer_sweep.sh <fixture-dir> [replicate]
fx=$1; rep=${2:-1}
run() { echo "== $*"; env "$@" python tools/bench_lamsa.py $fx --skip-reference --repeat 1 --replicate $rep --in-flight ${INFLIGHT:-8192} 2>&1 | grep -E "lamsa_b200\]|impl" | sed -e 's/"impl".*"stage_s"/"stage_s"/' | cut -c1-700; }
run LB2_NOOP=1
run LB2_HOST_THREADS=8
run LB2_MIN_BATCH=512 LB2_GATHER_US=100
run LB2_MIN_BATCH=8192 LB2_GATHER_US=1000
run LB2_FLUSH_TASKS=64
run LB2_FAST_ROWS=0
run LB2_FAST_ROWS=1024
