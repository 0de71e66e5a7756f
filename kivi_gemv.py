"""The reference's pybind extension module `kivi_gemv` (quant/csrc/pybind.cpp:5-8, imported by quant/matmul.py:6):
gemv_forward_cuda and gemv_forward_cuda_outer_dim, here bound to libkivi_b200.so through kivi_b200.kivi_gemv."""
from kivi_b200.kivi_gemv import gemv_forward_cuda, gemv_forward_cuda_outer_dim            # noqa: F401
