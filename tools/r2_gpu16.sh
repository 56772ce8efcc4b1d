#!/bin/bash
# usage: tools/r2_gpu16.sh TAG   one full ncu capture of a whole C3 step (128 pictures per launch) on the current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VC2_CODEC_SUBBATCH=1 timeout 900 ncu --set full --clock-control none --import-source on -s 28 -c 14 -f -o gpurun_out/$1_step \
  python tools/profile_step.py C3 1 128 > gpurun_out/$1_ncu_full.log 2>&1
tail -3 gpurun_out/$1_ncu_full.log
