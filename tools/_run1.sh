python tools/sweep.py 8 16 32 48 64 80 96 112 128 141 160 176 192 > gpurun_out/r02_sweep_small.jsonl 2>&1
python tools/sweep_configs.py > gpurun_out/r02_sweep_configs.jsonl 2>&1
python tools/latency_probe.py > gpurun_out/r02_latency.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:logl_onchip -s 2 -c 1 -o gpurun_out/r02c_onchip_N64 -f python tools/sweep.py 64 > gpurun_out/ncu_N64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logl_onchip -s 2 -c 1 -o gpurun_out/r02c_onchip_N141 -f python tools/sweep.py 141 > gpurun_out/ncu_N141.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logl_onchip -s 2 -c 1 -o gpurun_out/r02c_onchip_C4 -f python tools/sweep_configs.py "C4 SLSN" > gpurun_out/ncu_C4.log 2>&1
