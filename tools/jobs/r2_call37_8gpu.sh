# round 2, session 2: 8-GPU and 4-GPU bench lines of the shipped build
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 8 --steps 32 --warmup 4 > gpurun_out/r02_bench_line_8gpu.json 2> gpurun_out/r2f_bench_8gpu.err; echo "bench8 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2f_bench_8gpu.err | tail -3 | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_line_8gpu.json').read().strip().split('\n')[-1]); print({k: d[k] for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step','n_gpus','logits_allgather_ms')}, d['collective'][:70]); print({k: (v.get('value'), v.get('ms_per_step'), v.get('error')) for k,v in d['extra_configs'].items()})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29744 bench.py --gpus 4 --steps 32 --warmup 4 --no-extra > gpurun_out/r02_bench_line_4gpu.json 2> gpurun_out/r2f_bench_4gpu.err; echo "bench4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_line_4gpu.json').read().strip().split('\n')[-1]); print({k: d[k] for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step','n_gpus')}, d['collective'][:70])"
