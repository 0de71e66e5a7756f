"""Import-path compatibility with the reference checkout: `from models.llama_kivi import LlamaForCausalLM_KIVI`
(example.py:4, mem_spd_test.py:4) and `from models.mistral_kivi import MistralForCausalLM_KIVI` resolve to kivi_b200."""
