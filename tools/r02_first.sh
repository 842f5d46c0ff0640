#!/bin/bash
# first GPU pass over the on-chip shape: parity, microbenchmarks, throughput against the workspace kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_onchip.py -x -q > gpurun_out/t_onchip.log 2>&1; echo "pytest onchip rc=$?" | tee -a gpurun_out/t_onchip.log
./bench/lat > gpurun_out/lat.json 2>&1
timeout 600 python tools/sweep.py 32 64 96 128 141 160 192 224 > gpurun_out/sw_onchip.jsonl 2>&1
GSN_FORCE_SHAPE=legacy timeout 600 python tools/sweep.py 32 64 96 128 141 160 192 224 > gpurun_out/sw_legacy.jsonl 2>&1
timeout 900 python tools/sweep_configs.py > gpurun_out/cfg_onchip.jsonl 2>&1
for T in 1 2 4; do GSN_ONCHIP_T=$T timeout 300 python tools/sweep.py 32 64 96 > gpurun_out/sw_T$T.jsonl 2>&1; done
for T in 2 4 8 16; do GSN_ONCHIP_T=$T timeout 300 python tools/sweep.py 128 141 192 224 > gpurun_out/sw_T$T.b.jsonl 2>&1; done
