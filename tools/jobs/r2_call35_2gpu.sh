# round 2, session 2: 2-GPU bench line of the shipped build (peer-store id exchange inside the step graph) + the exchange tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_exchange_gpu.py -m gpu -q > gpurun_out/r2f_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests2.log; tail -3 gpurun_out/r2f_tests2.log | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/r02_bench_line_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/r2f_bench_2gpu.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_line_2gpu.json').read().strip().split('\n')[-1]); print({k: d[k] for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step','n_gpus')}, d['collective'][:70]); print({k: (v.get('value'), v.get('ms_per_step'), v.get('error')) for k,v in d['extra_configs'].items()})"
