python tools/sweep.py 8 16 32 48 64 80 96 112 128 141 160 176 192 > gpurun_out/r02_sweep_small.jsonl 2>&1
python tools/sweep_configs.py > gpurun_out/r02_sweep_configs.jsonl 2>&1
python tools/latency_probe.py > gpurun_out/r02_latency.jsonl 2>&1
python bench.py --steps 10 --warmup 3 --peaks > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
