#!/bin/bash
# usage: tools/r2_gpu8.sh TAG   full GPU parity suite, then the C3 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/$1.tests.txt
tools/r2_bench.sh $1 "C3"
