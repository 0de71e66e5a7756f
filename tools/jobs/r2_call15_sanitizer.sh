mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 0 --target-processes all --error-exitcode 7 python -m pytest tests/test_bgemv_gpu.py -m gpu -q -x -k "sv_shape or qk_shape" > gpurun_out/r2o_sanitizer_bgemv.log 2>&1; echo "bgemv memcheck rc=$?"; tail -5 gpurun_out/r2o_sanitizer_bgemv.log | cut -c1-200
timeout 1200 compute-sanitizer --tool memcheck --launch-timeout 0 --target-processes all --error-exitcode 7 python -m pytest tests/test_decode_gpu.py -m gpu -q -x -k "decode_steps_match_oracle and G-auto or import_tuple or capacity_guard or decode_with_mask" > gpurun_out/r2o_sanitizer_decode.log 2>&1; echo "decode memcheck rc=$?"; tail -5 gpurun_out/r2o_sanitizer_decode.log | cut -c1-200
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r2o_sanitizer_*.log
unset PYTORCH_NO_CUDA_MEMORY_CACHING
timeout 300 python tools/microbench.py --only-fused 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prefill cfg2', d['prefill_pack_ms'], d['prefill_pack_GBps'])"
