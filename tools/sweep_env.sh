#!/bin/bash
# usage: tools/sweep_env.sh  -- runs tools/profile_step.py C3 under a list of tuning environments (GPU box)
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" python tools/profile_step.py C3 5 8 2>&1 | grep -E "dwt|idwt" ; }
run VC2_DWT_PD=2
run VC2_DWT_PD=4
run VC2_DWT_PD=6
run VC2_DWT_PD=10
run VC2_DWT_PD=4 VC2_DWT_MIN_WARPS=14208
run VC2_DWT_PD=4 VC2_DWT_MIN_WARPS=28416
run VC2_DWT_PD=4 VC2_DWT_SEG_ROWS=64
run VC2_DWT_PD=4 VC2_DWT_SEG_ROWS=32
run VC2_DWT_PD=8 VC2_DWT_SEG_ROWS=32
