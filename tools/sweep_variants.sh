# time every tuning build in tools/variants/ on the cfg-2 layer shape (tools/microbench.py)
for so in tools/variants/libkivi_*.so; do
  KIVI_B200_LIB=$PWD/$so timeout 120 python tools/microbench.py 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$so', d['fused_decode_ms'], d['fused_decode_best_ms'])"
done
