# round 2, session 2, call 8: the adopted set (grid-constant params; q first and window logits in the copy group for G = 1) against the
# start of the session (nogc) and grid-constant alone (gc_only), same box; GPU suite on the adopted build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s2_tests5.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s2_tests5.log; tail -4 gpurun_out/r2s2_tests5.log | cut -c1-300
for rep in 1 2 3; do
for v in default nogc gc_only; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  if [ $rep = 1 ]; then timeout 300 python tools/ab_fused.py cfg2 cfg3 cfg4 b128 k4mha 2>/dev/null; else timeout 300 python tools/ab_fused.py cfg2 cfg3 2>/dev/null; fi
done; done 2>&1 | tee gpurun_out/r2s2_ab7.txt
