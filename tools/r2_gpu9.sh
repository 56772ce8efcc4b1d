#!/bin/bash
# usage: tools/r2_gpu9.sh TAG [pytest -k expression]   A/B probe of the current library against the baseline (bit exactness of 14 cases + stage times), then selected tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so > gpurun_out/$1.probe.txt 2>&1
echo "probe rc=$?: $(tail -1 gpurun_out/$1.probe.txt)"; grep -A9 "C3 DD137" gpurun_out/$1.probe.txt | head -10
if [ -n "$2" ]; then python -m pytest tests -q -m gpu -k "$2" 2>&1 | tail -8 | tee gpurun_out/$1.tests.txt; fi
