"""quant/matmul.py of the reference -> kivi_b200.matmul (cuda_bmm_fA_qB_outer, triton_bmm_fA_qB_outer)."""
from kivi_b200.matmul import cuda_bmm_fA_qB_outer, triton_bmm_fA_qB_outer            # noqa: F401
