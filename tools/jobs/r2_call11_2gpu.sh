mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_exchange_gpu.py tests/test_decode_gpu.py -m gpu -q -k "exchange or greedy or second_device" > gpurun_out/r2k_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests2.log; tail -15 gpurun_out/r2k_tests2.log | cut -c1-300
for mode in p2p nccl p2p nccl; do
  if [ $mode = nccl ]; then export KIVI_BENCH_COLLECTIVE=nccl; else unset KIVI_BENCH_COLLECTIVE; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 32 --warmup 4 --no-extra > gpurun_out/r2k_bench_2gpu_$mode.json 2> gpurun_out/r2k_bench_2gpu_$mode.err; echo "bench2 $mode rc=$?"; tail -2 gpurun_out/r2k_bench_2gpu_$mode.err | cut -c1-300
  python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_2gpu_$mode.json').read().strip().split('\n')[-1]); print('$mode', {k: d[k] for k in ('value','ms_per_step','median_ms_per_step','max_ms_per_step','n_gpus')}, d['collective'][:60])"
done
