python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -4
python tools/sweep.py 16 32 48 64 96 128 141 192 > gpurun_out/sw_r2.jsonl 2>&1
python tools/sweep_configs.py "C4 SLSN" C2 "C4 dynesty" >> gpurun_out/sw_r2.jsonl 2>&1
