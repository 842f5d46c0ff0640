python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -2
python tools/sweep.py 8 16 32 48 64 > gpurun_out/sw_stage.jsonl 2>&1
python tools/sweep_configs.py C1 "C4 SLSN" "C4 dynesty" >> gpurun_out/sw_stage.jsonl 2>&1
