# round 2, session 2, call 1: validate HEAD in the re-created container + A/B of KIVI_COMMIT_LATE (p.V cache updates after the
# first stages are in flight) and KIVI_PREFETCH_SV (finished q.K^T warps request the first V items of the p.V ranges into L2)
mkdir -p gpurun_out
for rep in 1 2; do
for v in default late pf2 pf4 late_pf2 late_pf4; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  if [ $rep = 1 ]; then timeout 300 python tools/ab_fused.py cfg2 cfg3 cfg4 b128 2>/dev/null; else timeout 300 python tools/ab_fused.py cfg2 cfg3 2>/dev/null; fi
done; done 2>&1 | tee gpurun_out/r2s2_ab1.txt
for v in tl tl_late tl_late_pf2; do
  echo "== timeline $v"
  KIVI_TL_OUT=gpurun_out/timeline_$v.npy KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so timeout 300 python tools/timeline.py 2>&1 | tail -16
done 2>&1 | tee gpurun_out/r2s2_timeline1.txt
export KIVI_B200_LIB=$PWD/tools/variants/libkivi_late_pf2.so
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r2s2_tests_late.log 2>&1; echo "late_pf2 tests rc=$?" >> gpurun_out/r2s2_tests_late.log; tail -3 gpurun_out/r2s2_tests_late.log | cut -c1-300
unset KIVI_B200_LIB
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests.log; tail -4 gpurun_out/r2s2_tests.log | cut -c1-300
