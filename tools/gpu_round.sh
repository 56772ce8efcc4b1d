#!/bin/bash
# usage (GPU box, one GPU): tools/gpu_round.sh <tag> [notests]
# parity tests, the bench line of both arms, the ncu launch list of the bench command and one full capture of a step
tag=${1:-r1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${tag}_tests.txt
fi
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("b200", d["value"], d["encode_fps"], d["decode_fps"], "e2e", d["e2e"]["value"], d["roofline"], d.get("cpu_baseline"))
print({k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()})
print(open("gpurun_out/${tag}_bench_reference.json").read()[:600])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-pictures 8 > gpurun_out/${tag}_ncu_bench.log 2>&1
VC2_CODEC_SUBBATCH=1 ncu --set full --clock-control none --import-source on -s 26 -c 13 -f -o gpurun_out/prof_${tag} \
  python tools/profile_step.py C3 1 128 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
