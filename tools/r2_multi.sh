#!/bin/bash
# usage (gpurun --gpus N -- tools/r2_multi.sh TAG N): BASELINE config 3 as a whole job on all N GPUs (240 frames, -G N, md5 against the
# reference's), then the N-rank bench line (weak scaling, end-to-end rates, synchronised PCIe probe)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader | head -8; nproc; numactl -H 2>/dev/null | head -4
VC2_SEQ_GPUS="$2" timeout 900 python -m pytest tests/test_gpu_sequence.py -q -m gpu -s 2>&1 | tail -6 | tee gpurun_out/$1.seq_tests.txt
cp gpurun_out/sequence_fps.json gpurun_out/$1_sequence_fps_g$2.json 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$1_bench_n$2.json 2> gpurun_out/$1_bench_n$2.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$1_bench_n$2.json"))
    e = d["e2e"]
    print("N=$2 value %.0f e2e %.0f frac_of_pcie_bound %.2f enc_only %.0f dec_only %.0f" % (d["value"], e["value"], e["frac_of_pcie_bound"], e["encode_only_fps"], e["decode_only_fps"]), e["pcie_pinned_copy_all_ranks"], "numa", e["numa_bound_ranks"])
except Exception as ex:
    print("bench failed", ex); print(open("gpurun_out/$1_bench_n$2.err").read()[-1500:])
PY
