for v in shipped accexp sqrtdiv both; do
  if [ $v = shipped ]; then unset GSN_LIB; else export GSN_LIB=$PWD/variants/$v.so; fi
  echo "variant $v" >> gpurun_out/arbiter_variants.jsonl
  SPANS=150 python tools/arbiter_check.py 192 512 >> gpurun_out/arbiter_variants.jsonl 2>> gpurun_out/arbiter_variants.err
done
