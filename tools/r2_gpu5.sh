#!/bin/bash
# usage: tools/r2_gpu5.sh TAG [notests]   probe A/B against the round-1 build (all stages), then the GPU parity suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${1:-r2_g9}
timeout 300 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so > $O.probe.txt 2>&1
cat $O.probe.txt | cut -c1-150
VC2_NARROW=0 timeout 300 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so > $O.probe_wide.txt 2>&1
grep -A9 "C3 DD137" $O.probe_wide.txt | cut -c1-150; tail -1 $O.probe_wide.txt
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $O.tests.txt
fi
