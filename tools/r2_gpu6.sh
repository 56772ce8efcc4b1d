#!/bin/bash
# usage: tools/r2_gpu6.sh TAG   several library builds side by side (probe, 32 C3 pictures) + LD / CBR bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/$1
timeout 600 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so tools/_probe/var/lib_b.so tools/_probe/var/lib_c.so tools/_probe/var/lib_d.so > $O.probe_var.txt 2>&1
grep -A9 "C3 DD137" $O.probe_var.txt | cut -c1-170; tail -1 $O.probe_var.txt
bash tools/r2_bench.sh $1 "C5"
