python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench3.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l3.log 2>&1; echo "ncu rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref3.json 2> gpurun_out/bench_r02_ref3.err; echo "ref rc=$?"
