mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -m gpu -q -k "import or full_size" > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log; tail -3 gpurun_out/r2c_tests.log
run() { # name, args...
  n=$1; shift
  timeout 200 python tools/microbench.py --only-fused "$@" 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v $n', round(d['fused_decode_ms'],5), round(d['fused_decode_best_ms'],5), round(d['fused_decode_noflush_ms'],5), round(d['fused_decode_GBps']))"
}
for v in default cw20 cw24 cw20u4 cw24u4; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  run cfg2
  run cfg3 --B 64 --H 32 --Hkv 8 --T 8192
  run cfg4 --B 16 --H 32 --Hkv 8 --T 32768 --bits 4 --g 64 --R 64
  run cfg5 --B 128 --T 4096
done > gpurun_out/r2c_variants.log 2>&1
cat gpurun_out/r2c_variants.log
