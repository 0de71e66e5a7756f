mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_decode_gpu.py -m gpu -q -x > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log; tail -4 gpurun_out/r2s_tests.log | cut -c1-300
for rep in 1 2; do
for v in pdl nopdl; do
  if [ $v = pdl ]; then unset KIVI_NO_GLUE_PDL; else export KIVI_NO_GLUE_PDL=1; fi
  timeout 600 python bench.py --steps 32 --warmup 4 --no-extra --no-cpu-baseline --no-reference-gpu > gpurun_out/r2s_bench_$v.json 2> gpurun_out/r2s_bench_$v.err; echo "bench $v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_$v.json').read().strip().split('\n')[-1]); print({k: round(d[k],3) for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step')}, round(d['roofline']['frac'],4), round(d['roofline']['launch_ms'],5), round(d['e2e']['value'],1))"
done; done
