"""quant/new_pack.py of the reference -> kivi_b200.new_pack (same names, same argument order)."""
from kivi_b200.new_pack import *            # noqa: F401,F403
from kivi_b200.new_pack import (pack_tensor, quant_and_pack_kcache, quant_and_pack_vcache,            # noqa: F401
                                triton_quantize_and_pack_along_last_dim, unpack_and_dequant_kcache,
                                unpack_and_dequant_vcache, unpack_tensor)
