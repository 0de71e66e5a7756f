set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log; tail -15 gpurun_out/r2b_tests.log
for v in default bprep3; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  for i in 1 2; do timeout 200 python tools/microbench.py --only-fused 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v cfg2', d['fused_decode_ms'], d['fused_decode_best_ms'], d['fused_decode_noflush_ms'])"; done
  timeout 200 python tools/microbench.py --only-fused --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v cfg3', d['fused_decode_ms'], d['fused_decode_best_ms'], d['fused_decode_noflush_ms'])"
done > gpurun_out/r2b_variants.log 2>&1
cat gpurun_out/r2b_variants.log
