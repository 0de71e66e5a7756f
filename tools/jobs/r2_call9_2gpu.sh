mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m pytest tests/test_decode_gpu.py -m gpu -q -k "second_device" > gpurun_out/r2i_tests2.log 2>&1; tail -3 gpurun_out/r2i_tests2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2i_bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_2gpu.json')); print({k: d[k] for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step','collective','logits_allgather_ms','n_gpus')}); print({k: (v.get('value'), v.get('ms_per_step'), v.get('error')) for k,v in d['extra_configs'].items()})"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2i_ref_2gpu.json 2> gpurun_out/r2i_ref_2gpu.err; echo "ref rc=$?"; head -c 600 gpurun_out/r2i_ref_2gpu.json
