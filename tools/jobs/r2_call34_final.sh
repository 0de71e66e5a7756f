# round 2, session 2: final measurements of the shipped build (tests, smoke, bench lines, ncu capture + launch list, microbenchmarks, timeline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log; tail -4 gpurun_out/r2f_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 32 --warmup 4 > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2>> gpurun_out/r02_bench.err; echo "ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'qk_kernel|sv_kernel' -s 4 -c 2 -f -o gpurun_out/r02_attn python tools/profile_fused.py > gpurun_out/r02_ncu.log 2>&1; tail -2 gpurun_out/r02_ncu.log
KIVI_PROFILE_STEPS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_bench_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02_bench_launches.csv
timeout 300 python tools/microbench.py --ref --out gpurun_out/r02_microbench.json > /dev/null 2>&1
timeout 300 python tools/microbench.py --B 16 --H 32 --Hkv 8 --T 32832 --bits 4 --g 64 --R 64 --out gpurun_out/r02_microbench_cfg4.json > /dev/null 2>&1
timeout 300 python tools/microbench.py --B 64 --H 32 --Hkv 8 --T 8192 --out gpurun_out/r02_microbench_cfg3.json > /dev/null 2>&1
timeout 300 python tools/ab_fused.py cfg2 cfg3 cfg4 b128 k4mha k4g128 k4gqa2 2>/dev/null | tee gpurun_out/r02_ab_final.txt
KIVI_TL_OUT=gpurun_out/timeline_final.npy KIVI_B200_LIB=$PWD/tools/variants/libkivi_tl.so timeout 300 python tools/timeline.py 2>&1 | tail -20 | tee gpurun_out/r02_timeline.txt
