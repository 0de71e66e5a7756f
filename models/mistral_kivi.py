"""models/mistral_kivi.py of the reference -> kivi_b200.llama_kivi (the Mistral hook is the Llama hook with grouped KV
heads; GQA is handled inside the kernels, without the repeat_kv_quant copies of models/mistral_kivi.py:58-67)."""
from kivi_b200.llama_kivi import (MistralFlashAttention_KIVI, MistralForCausalLM_KIVI, repeat_kv)            # noqa: F401
