# round 2, session 2, call 2: ticketed cache updates in the q.K^T tail (KIVI_COMMIT_IN_QK, the new default) vs the p.V-prologue
# version (nocq); 12 / 13 warps per CTA (two stages per warp for the G = 4 kernels, which get one at 16 warps); tests; timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s2_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests2.log; tail -4 gpurun_out/r2s2_tests2.log | cut -c1-300
for rep in 1 2 3; do
for v in default nocq cw13 cw12; do
  if [ $rep = 3 ] && [ $v != default ] && [ $v != nocq ]; then continue; fi
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  if [ $rep = 1 ]; then timeout 300 python tools/ab_fused.py cfg2 cfg3 cfg4 b128 2>/dev/null; else timeout 300 python tools/ab_fused.py cfg2 cfg3 2>/dev/null; fi
done; done 2>&1 | tee gpurun_out/r2s2_ab2.txt
echo "== timeline tl (ticketed cache updates in q.K^T)"
KIVI_TL_OUT=gpurun_out/timeline_tl_cq.npy KIVI_B200_LIB=$PWD/tools/variants/libkivi_tl.so timeout 300 python tools/timeline.py 2>&1 | tail -20 | tee gpurun_out/r2s2_timeline2.txt
unset KIVI_B200_LIB
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 600 compute-sanitizer --tool memcheck --launch-timeout 0 --target-processes all --error-exitcode 7 python -m pytest tests/test_decode_gpu.py -m gpu -q -x -k "decode_steps_match_oracle and G-auto or import_tuple or capacity_guard or decode_with_mask" > gpurun_out/r2s2_sanitizer_decode.log 2>&1; echo "decode memcheck rc=$?"; tail -3 gpurun_out/r2s2_sanitizer_decode.log | cut -c1-200
