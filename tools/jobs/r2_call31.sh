# round 2, session 2, call 7: the latency fixes read off the ncu source counters: q before the bulk of the first stages (KIVI_Q_FIRST),
# release-only p.V arrival (KIVI_REL_ARRIVE), window-item logits in the bulk-copy group (KIVI_WIN_LOGITS_BULK), __grid_constant__ params.
# default = all on; base = all three off; qf0 / ra0 / wl0 = one off.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s2_tests4.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests4.log; tail -4 gpurun_out/r2s2_tests4.log | cut -c1-300
for rep in 1 2 3; do
for v in default base qf0 ra0 wl0; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  if [ $rep = 1 ]; then timeout 300 python tools/ab_fused.py cfg2 cfg3 cfg4 b128 2>/dev/null; else timeout 300 python tools/ab_fused.py cfg2 cfg3 2>/dev/null; fi
done; done 2>&1 | tee gpurun_out/r2s2_ab6.txt
echo "== timeline"
KIVI_TL_OUT=gpurun_out/timeline_tl_s2.npy KIVI_B200_LIB=$PWD/tools/variants/libkivi_tl.so timeout 300 python tools/timeline.py 2>&1 | tail -20 | tee gpurun_out/r2s2_timeline3.txt
