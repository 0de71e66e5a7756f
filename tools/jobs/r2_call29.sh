# round 2, session 2, call 5: the build with 12-warp CTAs for the 4-bit G = 4 kernels: GPU suite, smoke, per-shape timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests3.log; tail -4 gpurun_out/r2s2_tests3.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for rep in 1 2; do timeout 300 python tools/ab_fused.py cfg2 cfg3 cfg4 b128 k4mha k4g128 k4gqa2 2>/dev/null; done | tee gpurun_out/r2s2_ab5.txt
