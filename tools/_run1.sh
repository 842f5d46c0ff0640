timeout 600 python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -4
timeout 300 python tools/sweep.py 80 96 112 128 141 160 176 192 > gpurun_out/sw_early.jsonl 2>&1
timeout 300 python tools/sweep_configs.py C2 "C3 real" "C4 synthetic" >> gpurun_out/sw_early.jsonl 2>&1
timeout 300 python tools/latency_probe.py > gpurun_out/lat_early.jsonl 2>&1
