mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py -m gpu -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2n_tests.log; tail -6 gpurun_out/r2n_tests.log | cut -c1-300
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k.startswith('fused_decode_ms') or k.startswith('fused_decode_graph')})"; }
for v in selfin default selfin default; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  timeout 200 python tools/microbench.py --only-fused 2>/dev/null | tail -1 | show fused_${v}_cfg2
  timeout 200 python tools/microbench.py --only-fused --B 64 --H 32 --Hkv 8 --T 8192 2>/dev/null | tail -1 | show fused_${v}_cfg3
  timeout 200 python tools/microbench.py --only-fused --B 16 --H 32 --Hkv 8 --T 32768 --bits 4 --g 64 --R 64 2>/dev/null | tail -1 | show fused_${v}_cfg4
  timeout 200 python tools/microbench.py --only-fused --B 128 2>/dev/null | tail -1 | show fused_${v}_cfg5s
done
for v in selfin default; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  timeout 600 python bench.py --steps 32 --warmup 4 --no-extra --no-cpu-baseline --no-reference-gpu > gpurun_out/r2n_bench_$v.json 2> gpurun_out/r2n_bench_$v.err; echo "bench $v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench_$v.json').read().strip().split('\n')[-1]); print('$v', {k: round(d[k],3) for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step')}, round(d['roofline']['frac'],4), round(d['roofline']['launch_ms'],5), round(d['e2e']['value'],1))"
done
