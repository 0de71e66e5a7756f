mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bgemv_gpu.py tests/test_decode_gpu.py tests/test_model_gpu.py -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log; tail -5 gpurun_out/r2e_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k.startswith('ours_q') or k.startswith('ours_s') or k.startswith('fused_decode_ms') or k.startswith('fused_decode_GB')})"; }
timeout 300 python tools/microbench.py 2>/dev/null | tail -1 | show mma_cfg2
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py 2>/dev/null | tail -1 | show simt_cfg2
timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show mma_cfg4b
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show simt_cfg4b
for v in uniform default uniform default; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  timeout 200 python tools/microbench.py --only-fused 2>/dev/null | tail -1 | show fused_${v}_cfg2
  timeout 200 python tools/microbench.py --only-fused --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | show fused_${v}_cfg3
  timeout 200 python tools/microbench.py --only-fused --B 16 --H 32 --Hkv 8 --T 32768 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show fused_${v}_cfg4
done
for v in tlu tl; do
  echo "== timeline $v"; KIVI_TL_OUT=gpurun_out/timeline_r2_$v.npy KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so timeout 200 python tools/timeline.py 2>&1 | tail -16
done
