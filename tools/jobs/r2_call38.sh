# round 2, session 2, last call: exp(x - m) as ex2(fma(x, log2 e, -m log2 e)) (KIVI_EXP_FMA): GPU suite on the variant + A/B
mkdir -p gpurun_out
export KIVI_B200_LIB=$PWD/tools/variants/libkivi_expfma.so
timeout 100 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r2s2_tests8.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests8.log; tail -3 gpurun_out/r2s2_tests8.log | cut -c1-200
for v in expfma exp0 expfma exp0; do
  export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so
  timeout 60 python tools/ab_fused.py cfg2 cfg3 2>/dev/null
done | tee gpurun_out/r2s2_ab10.txt
