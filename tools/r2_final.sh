#!/bin/bash
# usage: tools/r2_final.sh TAG   the round's record on one B200: bench lines of every BASELINE configuration, the reference arm, the
# ncu launch list of the bench command and one full ncu capture of a whole C3 step (128 pictures per launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/r2_bench.sh $1 "C3 C1 C2 C4a C4b C5"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/$1_bench_reference.json 2> gpurun_out/$1_bench_reference.err; head -c 700 gpurun_out/$1_bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/$1_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-pictures 8 > gpurun_out/$1_ncu_bench.log 2>&1
VC2_CODEC_SUBBATCH=1 timeout 900 ncu --set full --clock-control none --import-source on -s 28 -c 14 -f -o gpurun_out/$1_step \
  python tools/profile_step.py C3 1 128 > gpurun_out/$1_ncu_full.log 2>&1
tail -12 gpurun_out/$1_ncu_full.log
ls -la gpurun_out | grep $1
