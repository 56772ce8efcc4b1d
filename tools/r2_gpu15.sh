#!/bin/bash
# usage: tools/r2_gpu15.sh TAG   full ncu capture (with source) of the HQ_CBR rate-control launch, 64 C2 pictures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VC2_CODEC_SUBBATCH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:${K:-hq_pack_kernel} -s ${SKIP:-4} -c 1 -f -o gpurun_out/$1_search python tools/profile_step.py C2 1 64 > gpurun_out/$1_search.log 2>&1
tail -3 gpurun_out/$1_search.log
