mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log; tail -4 gpurun_out/r2r_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k in ('fused_decode_ms','fused_decode_graph_ms','fused_decode_graph_GBps')})"; }
for rep in 1 2; do
for v in -1 1 2 3 4; do
  export KIVI_FUSED_GROUPS=$v
  timeout 300 python tools/microbench.py --only-fused 2>/dev/null | tail -1 | show g${v}_cfg2
  timeout 300 python tools/microbench.py --only-fused --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | show g${v}_cfg3
  timeout 300 python tools/microbench.py --only-fused --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show g${v}_cfg4
done; done
unset KIVI_FUSED_GROUPS
timeout 600 python bench.py --steps 32 --warmup 4 --no-extra --no-cpu-baseline --no-reference-gpu > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().split('\n')[-1]); print({k: round(d[k],3) for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step')}, round(d['roofline']['frac'],4), round(d['roofline']['launch_ms'],5), round(d['e2e']['value'],1))"
