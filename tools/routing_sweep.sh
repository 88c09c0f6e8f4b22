#!/bin/bash
# A/B of the kernel-routing thresholds (dp_pack.h: pick_variant) on the C2 microbenchmark: one JSON line per setting.
n=${1:-200000}
run() { env "$@" python tools/kernel_probe.py $n; }
run LB2_NOOP=1
run LB2_SUB16_NP4_MIN_EXT=375
run LB2_SUB16_NP4_MIN_EXT=300
run LB2_SUB_NP4_MIN_EXT=120
run LB2_SUB_NP4_MIN_EXT=200
run LB2_SUBWARP_MAX_GLB=300
run LB2_SUBWARP_MAX_GLB=150
run LB2_NP4_MIN=150
run LB2_SUB16_NP4_MIN_GLB=200 LB2_SUB16_NP4_MAX_GLB=410 LB2_SUBWARP_MAX_GLB=410
