#!/bin/bash
# usage: tools/r2_prof.sh TAG "regex:skip regex:skip ..."   source-level counters of single launches (probe, 8 C3 pictures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in $2; do
  k=${spec%%:*}; sk=${spec##*:}
  VC2_CODEC_SUBBATCH=1 timeout 900 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:$k -s $sk -c 1 -f -o gpurun_out/$1_$k tools/_probe/pack_probe 8 vc2_reference_b200/libvc2b200.so > gpurun_out/$1_$k.log 2>&1
done
ls gpurun_out | grep $1
