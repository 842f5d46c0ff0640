timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/sweep.py 32 48 64 96 128 141 192 512 > gpurun_out/sw_rsq.jsonl 2>&1
timeout 300 python tools/sweep_configs.py "C4 SLSN" C2 "C3 real" >> gpurun_out/sw_rsq.jsonl 2>&1
