# decode attention per layer at the BASELINE.json configs (tools/microbench.py --only-fused)
echo cfg2; timeout 200 python tools/microbench.py --only-fused | tail -1
echo cfg3; timeout 200 python tools/microbench.py --only-fused --B 64 --H 32 --Hkv 8 --T 8192 | tail -1
echo cfg4; timeout 200 python tools/microbench.py --only-fused --B 16 --H 32 --Hkv 8 --T 32768 --bits 4 --g 64 --R 64 | tail -1
echo cfg5-shard; timeout 200 python tools/microbench.py --only-fused --B 128 --T 4096 | tail -1
