# round 2, session 2, call 4: fewer warps per CTA for the 4-bit K kernels: 12 / 11 / 10
mkdir -p gpurun_out
for rep in 1 2; do
for v in cw12 cw11 cw10; do
  export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so
  timeout 300 python tools/ab_fused.py cfg4 k4mha k4g128 k4gqa2 2>/dev/null
done; done 2>&1 | tee gpurun_out/r2s2_ab4.txt
