mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bgemv_gpu.py tests/test_reference_cases_gpu.py tests/test_model_gpu.py -m gpu -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log; tail -25 gpurun_out/r2d_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k.startswith('ours_') or k.startswith('ref_')})"; }
timeout 300 python tools/microbench.py --ref 2>/dev/null | tail -1 | show mma_cfg2
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py 2>/dev/null | tail -1 | show simt_cfg2
timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32768 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show mma_cfg4
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32768 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show simt_cfg4
timeout 300 python tools/microbench.py --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | show mma_cfg3
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | show simt_cfg3
