#!/bin/bash
# usage: tools/r2_gpu14.sh TAG   slice index walk: priority streams on / off, 16 KB / 8 KB ring segments (device-resident C3 bench line each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-pictures 16 2>gpurun_out/$TAG.err_$label.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', 'rt=%.0f enc=%.0f dec=%.0f e2e=%.0f' % (d['value'], d['encode_fps'], d['decode_fps'], d['e2e']['value']))
" | tee -a gpurun_out/$TAG.index_ab.txt
}
TAG=$1
run prio0_seg16 VC2_INDEX_PRIO=0
run prio1_seg16 VC2_INDEX_PRIO=1
cp vc2_reference_b200/libvc2b200.so /tmp/lib_keep.so
cp tools/_probe/var/lib_seg8.so vc2_reference_b200/libvc2b200.so
run prio0_seg8 VC2_INDEX_PRIO=0
run prio1_seg8 VC2_INDEX_PRIO=1
cp /tmp/lib_keep.so vc2_reference_b200/libvc2b200.so
timeout 200 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so tools/_probe/var/lib_seg8.so > gpurun_out/$TAG.probe.txt 2>&1; tail -1 gpurun_out/$TAG.probe.txt; grep -A9 "C3 DD137" gpurun_out/$TAG.probe.txt | grep index
