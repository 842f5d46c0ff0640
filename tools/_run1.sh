python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -2
python tools/sweep.py 8 16 32 48 64 80 96 112 128 141 160 176 192 > gpurun_out/r02_sweep_small.jsonl 2>&1
python tools/sweep_configs.py > gpurun_out/r02_sweep_configs.jsonl 2>&1
python tools/latency_probe.py > gpurun_out/r02_latency.jsonl 2>&1
for c in 12 20; do :; done
