#!/bin/bash
# usage (GPU box): tools/variant_bench.sh <label> [env...] -- one short device-resident bench line, summarised
label=$1; shift
env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $VB_ARGS 2>gpurun_out/err_$label.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', 'rt=%.0f enc=%.0f dec=%.0f e2e=%.0f' % (d['value'], d['encode_fps'], d['decode_fps'], d['e2e']['value']), {k:round(v['ms_per_step'],3) for k,v in d['stages'].items()})
"
