#!/bin/bash
# usage: tools/r2_gpu10.sh TAG   A/B of the variant libraries in tools/_probe/var against the baseline (32 C3 pictures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so tools/_probe/var/lib_*.so > gpurun_out/$1.probe_var.txt 2>&1
echo "rc=$? $(ls tools/_probe/var/lib_*.so | tr '\n' ' ')"; grep -B1 -A9 "C3 DD137" gpurun_out/$1.probe_var.txt | grep -E "idwt|SAME|DIFF|status"; tail -1 gpurun_out/$1.probe_var.txt
