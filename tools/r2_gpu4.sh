#!/bin/bash
# probe timing + the whole GPU parity suite with the tile kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${1:-r2_g8}
bash tools/r2_gpu3.sh ${1:-r2_g8} "3 5"
VC2_DWT_TILE=3 timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $O.tests_tile.txt
