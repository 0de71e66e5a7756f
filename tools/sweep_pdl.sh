for v in 0 1; do
  if [ $v = 1 ]; then export KIVI_NO_PDL=1; else unset KIVI_NO_PDL; fi
  timeout 120 python tools/microbench.py 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no_pdl=$v', d['fused_decode_ms'], d['fused_decode_best_ms'], d['fused_decode_noflush_ms'])"
done
