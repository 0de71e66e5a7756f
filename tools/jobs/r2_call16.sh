mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log; tail -5 gpurun_out/r2p_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k.startswith('ours_pack') or k.startswith('prefill') or k.startswith('fused_decode_graph')})"; }
timeout 300 python tools/microbench.py 2>/dev/null | tail -1 | show cfg2
timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show cfg4b
timeout 600 python bench.py --steps 32 --warmup 4 --no-extra --no-cpu-baseline --no-reference-gpu > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().split('\n')[-1]); print({k: round(d[k],3) for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step')}, round(d['roofline']['frac'],4), round(d['roofline']['launch_ms'],5), round(d['e2e']['value'],1))"
