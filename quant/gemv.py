"""quant/gemv.py of the reference -> kivi_b200.gemv (gemv_fwd, dequant_weight, dequant_weight_outer)."""
from kivi_b200.gemv import dequant_weight, dequant_weight_outer, gemv_fwd            # noqa: F401
