#!/bin/bash
# usage: tools/r2_gpu7.sh TAG   forward level 0 with TMA staging against plain loads (probe: bit exactness and stage times)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/$1
for t in 1 0; do
  VC2_DWT_TMA=$t timeout 120 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so > $O.probe_tma$t.txt 2>&1
  echo "TMA=$t rc=$?: $(tail -1 $O.probe_tma$t.txt)"; grep -A3 "C3 DD137" $O.probe_tma$t.txt | grep "dwt"
done
