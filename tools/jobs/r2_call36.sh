# round 2, session 2: deferred p.V arrival (KIVI_DEFER_ARRIVE, default build) vs immediate (da0)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s2_tests7.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests7.log; tail -4 gpurun_out/r2s2_tests7.log | cut -c1-300
for rep in 1 2 3; do
for v in default da0; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  if [ $rep = 1 ]; then timeout 300 python tools/ab_fused.py cfg2 b128 k4mha 2>/dev/null; else timeout 300 python tools/ab_fused.py cfg2 b128 2>/dev/null; fi
done; done 2>&1 | tee gpurun_out/r2s2_ab9.txt
