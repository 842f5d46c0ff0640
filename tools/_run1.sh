python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --peaks > gpurun_out/bench_r02_s2.json 2> gpurun_out/bench_r02_s2.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
