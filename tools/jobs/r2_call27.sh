# round 2, session 2, call 3: warps per CTA for the 4-bit K kernels (cfg 4 ran 4.5 % faster at 12 warps in call 2): 16 / 14 / 12 on four 4-bit shapes
mkdir -p gpurun_out
for rep in 1 2; do
for v in default cw14 cw12; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  timeout 300 python tools/ab_fused.py cfg4 k4mha k4g128 k4gqa2 2>/dev/null
done; done 2>&1 | tee gpurun_out/r2s2_ab3.txt
