#!/bin/bash
# round 2, GPU call 1: tile DWT kernels - parity vs the reference, A/B timing vs the streaming kernels, instruction counts
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_g1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O.smi.txt
tools/_probe/ubench > $O.ubench.txt 2>&1
for t in 1 2; do
  VC2_DWT_TILE=$t timeout 600 python -m pytest tests/test_gpu_library.py -x -q -m gpu -k "dwt" 2>&1 | tail -8 > $O.test_lib_tile$t.txt
done
VC2_DWT_TILE=1 timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -8 > $O.test_codec_tile1.txt
for t in 0 1 2; do
  VC2_DWT_TILE=$t timeout 300 tools/_probe/pack_probe 32 tools/_probe/baseline.so vc2_reference_b200/libvc2b200.so > $O.probe_tile$t.txt 2>&1
done
M=smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum
for t in 1 2; do
  VC2_DWT_TILE=$t VC2_CODEC_SUBBATCH=1 timeout 600 ncu --metrics $M --clock-control none -k regex:dwt_ -s 20 -c 10 --csv --log-file $O.ncu_tile$t.csv tools/_probe/pack_probe 8 vc2_reference_b200/libvc2b200.so vc2_reference_b200/libvc2b200.so > $O.ncu_tile$t.log 2>&1
done
VC2_DWT_TILE=0 VC2_CODEC_SUBBATCH=1 timeout 600 ncu --metrics $M --clock-control none -k regex:dwt_ -s 20 -c 10 --csv --log-file $O.ncu_tile0.csv tools/_probe/pack_probe 8 vc2_reference_b200/libvc2b200.so vc2_reference_b200/libvc2b200.so > $O.ncu_tile0.log 2>&1
tail -3 $O.test_*.txt; grep -A12 "C3 DD137" $O.probe_tile1.txt; tail -1 $O.probe_tile*.txt
