for v in 10_6_48_c1 9_7_48_c1 10_6_48_c0; do
timeout -s KILL 120 ./bench/ozaki_syrk_$v 64 > gpurun_out/ozaki_syrk_$v.json 2> gpurun_out/ozaki_syrk_$v.err; echo "ozaki $v rc=$?"
cat gpurun_out/ozaki_syrk_$v.json gpurun_out/ozaki_syrk_$v.err | head -5
done
