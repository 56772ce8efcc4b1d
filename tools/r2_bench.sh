#!/bin/bash
# usage: tools/r2_bench.sh TAG "configs"   bench lines of the given BASELINE configurations (default C3) into gpurun_out/TAG_bench_<cfg>.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in ${2:-C3}; do
  extra=""
  [ "$c" != "C3" ] && extra="--steps 10 --no-cpu-baseline"
  timeout 900 python bench.py --config $c $extra > gpurun_out/$1_bench_$c.json 2> gpurun_out/$1_bench_$c.err || tail -5 gpurun_out/$1_bench_$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$1_bench_$c.json"))
    print("$c", "value %.0f enc %.0f dec %.0f e2e %.0f" % (d["value"], d["encode_fps"], d["decode_fps"], d["e2e"]["value"]), "pipe", {k: round(v, 3) for k, v in d["pipeline_roofline"].items()})
    print("   ", {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}, d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d.get("cpu_baseline"))
except Exception as e:
    print("$c failed", e)
PY
done
