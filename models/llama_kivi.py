"""models/llama_kivi.py of the reference -> kivi_b200.llama_kivi."""
from kivi_b200.llama_kivi import (LlamaAttention_KIVI, LlamaDecoderLayer_KIVI, LlamaFlashAttention_KIVI,            # noqa: F401
                                  LlamaForCausalLM_KIVI, LlamaModel_KIVI, repeat_kv)
