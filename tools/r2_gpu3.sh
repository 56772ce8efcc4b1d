#!/bin/bash
# usage: tools/r2_gpu3.sh TAG "cfgs to time" [cfg to profile]
# A/B timing of the tile kernels (probe) + source-level counters of one configuration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${1:-r2_g3}
for t in $2; do
  VC2_DWT_TILE=$t timeout 300 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so > $O.probe_tile$t.txt 2>&1
  echo "cfg $t: $(tail -1 $O.probe_tile$t.txt)"; grep -A8 "C3 DD137" $O.probe_tile$t.txt | grep "dwt"
done
T=$3
if [ -n "$T" ]; then
VC2_DWT_TILE=$T VC2_CODEC_SUBBATCH=1 timeout 900 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --import-source on --clock-control none -k regex:dwt_tile_fwd -s 4 -c 1 -f -o ${O}_fwd$T tools/_probe/pack_probe 8 vc2_reference_b200/libvc2b200.so > ${O}_fwd$T.log 2>&1
VC2_DWT_TILE=$T VC2_CODEC_SUBBATCH=1 timeout 900 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --import-source on --clock-control none -k regex:dwt_tile_inv -s 7 -c 1 -f -o ${O}_inv$T tools/_probe/pack_probe 8 vc2_reference_b200/libvc2b200.so > ${O}_inv$T.log 2>&1
fi
