#!/bin/bash
# usage: tools/r2_gpu11.sh TAG   C++ mirror test + -n 3/4 command lines, then one full ncu capture of the HQ_CBR rate-control launch (C2, 64 pictures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -k "cpp_library_mirror or N4_ or N3_ or ld_quantise" 2>&1 | tail -6 | tee gpurun_out/$1.tests.txt
VC2_CODEC_SUBBATCH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hq_pack_kernel -s 4 -c 1 -f -o gpurun_out/$1_search python tools/profile_step.py C2 1 64 > gpurun_out/$1_search.log 2>&1
tail -12 gpurun_out/$1_search.log
