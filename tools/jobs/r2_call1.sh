set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log; tail -5 gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 16 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
rm -f gpurun_out/r2a_mem_spd.jsonl
timeout 600 python tools/mem_spd_test.py --out gpurun_out/r2a_mem_spd.jsonl > gpurun_out/r2a_memspd.log 2>&1; echo "memspd rc=$?"
timeout 900 python tools/mem_spd_test.py --fp16-baseline --out gpurun_out/r2a_mem_spd.jsonl >> gpurun_out/r2a_memspd.log 2>&1; echo "memspd16 rc=$?"
tail -4 gpurun_out/r2a_memspd.log
timeout 300 python tools/microbench.py --ref --out gpurun_out/r2a_micro.json > gpurun_out/r2a_micro.log 2>&1; echo "micro rc=$?"
