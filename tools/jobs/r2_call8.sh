mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py tests/test_bgemv_gpu.py -m gpu -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log; tail -12 gpurun_out/r2h_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k.startswith('fused_decode_ms') or k.startswith('fused_decode_graph') or k.startswith('ours_q') or k.startswith('ours_s')})"; }
timeout 200 python tools/microbench.py 2>/dev/null | tail -1 | show cfg2
timeout 200 python tools/microbench.py --only-fused --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | show cfg3
timeout 200 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show cfg4b
timeout 600 python bench.py --steps 32 --warmup 4 --no-extra --no-cpu-baseline --no-reference-gpu > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print({k: round(d[k],3) for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step')}, round(d['roofline']['frac'],4), round(d['roofline']['launch_ms'],5), round(d['e2e']['value'],1))"
