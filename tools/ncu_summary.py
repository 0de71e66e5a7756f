"""Summarise an ncu report of the decode-attention kernels into profiles/ (text + json).

    python tools/ncu_summary.py gpurun_out/t24_attn.ncu-rep profiles/r01_attention_ncu

Reads the report with `ncu -i ... --page raw --csv` / `--page source --csv` (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys
from collections import Counter

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__sass_inst_executed_op_tma_ld.sum"]
text, js = [], {"report": rep, "kernels": []}
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"]
    text.append(f"== {name}")
    k = {"kernel": name}
    for key in KEYS:
        if key in d:
            text.append(f"{key:75s} {d[key]:>18s} {u[key]}")
            try:
                k[key] = float(d[key].replace(",", ""))
            except ValueError:
                pass
    stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v.replace(",", ""))
              for h, v in d.items() if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")}
    top = sorted(stalls.items(), key=lambda x: -x[1])[:10]
    text.append("warp-cycles stalled per issued instruction: " + ", ".join(f"{a} {b:.2f}" for a, b in top))
    k["stalls_per_issue"] = dict(top)
    js["kernels"].append(k)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
idx = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
seen = set()
for n, start in enumerate(idx):
    name = srows[start][1]
    if name in seen:
        continue
    seen.add(name)
    h = srows[start + 1]
    end = idx[n + 1] if n + 1 < len(idx) else len(srows)
    isrc, ins, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ops, c, cs, ci = Counter(), Counter(), Counter(), Counter()
    for r in srows[start + 2:end]:
        try:
            ne, s = int(r[ins]), int(r[ismp])
        except (ValueError, IndexError):
            continue
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += ne
        c[ne] += ne
        cs[ne] += s
        ci[ne] += 1
    tot, tots = sum(c.values()), sum(cs.values())
    text.append(f"== {name}: {tot / 1e6:.2f} M warp instructions, {tots} stall samples")
    text.append("opcode mix: " + ", ".join(f"{o} {100 * v / tot:.1f}%" for o, v in ops.most_common(12)))
    text.append("instructions grouped by execution count (126976 / 63488 = chunk loop, 31744 = per packed block, ...):")
    for b, v in sorted(c.items(), key=lambda x: -cs[x[0]])[:8]:
        text.append(f"   executed {b:7d}x: {ci[b]:4d} instructions, {v / 1e6:6.2f} M ({100 * v / tot:4.1f} %), {100 * cs[b] / tots:4.1f} % of the stall samples")
js["dram_bytes_per_launch"] = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in js["kernels"][:2]) * 1e6
open(out + "_summary.txt", "w").write("\n".join(text) + "\n")
json.dump(js, open(out + ".json", "w"), indent=1)
print("\n".join(text))
