"""SASS evidence for the hot kernels of libkivi_b200.so (runs on the CPU box: cuobjdump only).

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt

Per kernel: instruction count and the mnemonics that prove what the source claims -- UBLKCP (cp.async.bulk = the TMA engine),
SYNCS (mbarrier), HMMA (mma.sync), LDSM (ldmatrix), LDGSTS (cp.async), LOP3 / PRMT / SHF (the unpack), LDL / STL (local
memory = spills or by-reference structs), and the absence of UTC*MMA / LDTM / UTMALDG (tcgen05 / tensor-map TMA are NOT used:
a GEMV over packed codes has no fp16 operand in canonical layout to feed them)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "kivi_b200", "csrc", "libkivi_b200.so")
WATCH = ["UBLKCP", "SYNCS", "HMMA", "LDSM", "LDGSTS", "LOP3", "PRMT", "SHF", "IMAD", "MUFU", "LDS", "STS", "LDG", "STG", "ATOMG", "RED",
         "LDL", "STL", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "BAR", "ERRBAR", "CCTL"]
KERNELS = [r"qk_kernel<2, 1, 32, 16>", r"sv_kernel<2, 2, 1, 32, 16>", r"qk_kernel<2, 4, 32, 16>", r"sv_kernel<2, 2, 4, 32, 16>",
           r"qk_kernel<4, 4, 64, 12>", r"sv_kernel<4, 4, 4, 64, 12>", r"bgm::wide_kernel<2, 1, 32>", r"bgm::tall_kernel<2, 1, 32>",
           r"pack_lastdim_kernel<2", r"block_prefill_kernel<2, true>"]

def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", out)[1:]
    print(f"# {os.path.relpath(SO, ROOT)}: {len(blocks)} kernels; counts are STATIC instructions (all code paths)")
    for pat in KERNELS:
        for name, body in zip(names, blocks):
            if re.search(pat, name):
                ops = collections.Counter()
                for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", body, flags=re.M):
                    ops[m.group(1)] += 1
                total = sum(ops.values())
                print(f"\n{name}\n  instructions {total}: " + ", ".join(f"{k} {ops[k]}" for k in WATCH if ops[k]))
                missing = [k for k in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG") if not ops[k]]
                print("  absent: " + ", ".join(missing))
                break
    return 0


if __name__ == "__main__":
    sys.exit(main())
