mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -q -k "decode_steps or full_size or import" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log; tail -5 gpurun_out/r2f_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k.startswith('ours_q') or k.startswith('ours_s') or k.startswith('ref_') or k.startswith('fused_decode_ms') or k.startswith('fused_decode_graph')})"; }
timeout 300 python tools/microbench.py --ref 2>/dev/null | tail -1 | show mma_cfg2
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py 2>/dev/null | tail -1 | show simt_cfg2
timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show mma_cfg4b
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show simt_cfg4b
timeout 300 python tools/microbench.py --B 128 2>/dev/null | tail -1 | show mma_cfg5
KIVI_NO_MMA_GEMV=1 timeout 300 python tools/microbench.py --B 128 2>/dev/null | tail -1 | show simt_cfg5
timeout 600 python bench.py --steps 32 --warmup 4 --no-extra > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print({k: d[k] for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step')}, d['roofline']['frac'], d['roofline']['launch_ms'], d['e2e']['value'])"
