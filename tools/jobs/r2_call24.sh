mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2w_tests.log; tail -3 gpurun_out/r2w_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k in ('fused_decode_ms','fused_decode_graph_ms','fused_decode_graph_GBps')})"; }
for rep in 1 2 3; do
for v in default serial; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  timeout 300 python tools/microbench.py --only-fused 2>/dev/null | tail -1 | show ${v}_cfg2
  timeout 300 python tools/microbench.py --only-fused --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | show ${v}_cfg3
done; done
