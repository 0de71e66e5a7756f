# round 2, session 2, call 9: inputs of the cache update as the first item of the stage queue (KIVI_COMMIT_ASYNC, default build) vs plain loads (as0)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s2_tests6.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests6.log; tail -4 gpurun_out/r2s2_tests6.log | cut -c1-300
for rep in 1 2 3; do
for v in default as0; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  if [ $rep = 1 ]; then timeout 300 python tools/ab_fused.py cfg2 cfg3 cfg4 b128 k4mha 2>/dev/null; else timeout 300 python tools/ab_fused.py cfg2 b128 2>/dev/null; fi
done; done 2>&1 | tee gpurun_out/r2s2_ab8.txt
echo "== timeline"
KIVI_TL_OUT=gpurun_out/timeline_tl_s2b.npy KIVI_B200_LIB=$PWD/tools/variants/libkivi_tl.so timeout 300 python tools/timeline.py 2>&1 | tail -20 | tee gpurun_out/r2s2_timeline4.txt
