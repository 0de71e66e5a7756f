"""CPU oracle for the KIVI decode hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package.  The product package
`kivi_b200` never does (tests/test_boundary.py greps for it).

Layout:
  kivi_oracle.c   plain-C restatement of pack / dequant / bgemv / softmax (fp16-exact)
  ref.py          numpy wrappers + the restated attention hook (decode step, 9-tuple cache)
  fake_quant.py   torch-CPU restatement of models/utils_quant.py fake-quant (CPU baseline)
  build.py        gcc recipe for kivi_oracle.c -> oracle/_build/libkivi_oracle.so
  build_ref.py    nvcc recipe for the UNMODIFIED reference extension -> oracle/_ref/kivi_gemv.so
"""
