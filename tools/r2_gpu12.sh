#!/bin/bash
# usage: tools/r2_gpu12.sh TAG "configs" [pytest -k expression]   probe (bit exactness, 14 cases), bench lines of the given configs, selected tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so > gpurun_out/$1.probe.txt 2>&1
echo "probe rc=$?: $(tail -1 gpurun_out/$1.probe.txt)"; grep -c SAME gpurun_out/$1.probe.txt; grep -A9 "C3 DD137" gpurun_out/$1.probe.txt | tail -8
tools/r2_bench.sh $1 "$2"
if [ -n "$3" ]; then python -m pytest tests -q -m gpu -k "$3" 2>&1 | tail -6 | tee gpurun_out/$1.tests.txt; fi
