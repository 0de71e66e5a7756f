mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k: round(v,4) for k,v in d.items() if k.startswith('fused_decode')})"; }
for rep in 1 2 3; do
for v in default u8; do
  if [ $v = default ]; then unset KIVI_B200_LIB; else export KIVI_B200_LIB=$PWD/tools/variants/libkivi_$v.so; fi
  timeout 300 python tools/microbench.py --only-fused 2>/dev/null | tail -1 | show ${v}_cfg2
  timeout 300 python tools/microbench.py --only-fused --B 128 2>/dev/null | tail -1 | show ${v}_b128
done; done
