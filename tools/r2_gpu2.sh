#!/bin/bash
# source-level instruction counts of the tile kernels (level 0, forward and inverse)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_g2
T=${1:-1}
VC2_DWT_TILE=$T VC2_CODEC_SUBBATCH=1 timeout 900 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --import-source on --clock-control none -k regex:dwt_tile_fwd -s 4 -c 1 -f -o ${O}_fwd$T tools/_probe/pack_probe 8 vc2_reference_b200/libvc2b200.so > ${O}_fwd$T.log 2>&1
VC2_DWT_TILE=$T VC2_CODEC_SUBBATCH=1 timeout 900 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --import-source on --clock-control none -k regex:dwt_tile_inv -s 7 -c 1 -f -o ${O}_inv$T tools/_probe/pack_probe 8 vc2_reference_b200/libvc2b200.so > ${O}_inv$T.log 2>&1
ls -la gpurun_out | tail -5
