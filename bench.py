#!/usr/bin/env python
"""bench.py -- decode tokens/s of Llama-2-7B with the KIVI (K2V2 g32 R128) cache on B200(s).

    python bench.py --gpus N --steps K --warmup W            # this repo (libkivi_b200 fused decode)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU fake-quant path

Metric (BASELINE.json): decode tokens/sec @ Llama-2-7B bs32 seq4k K2V2 g32 R128.  A "step" is one
decode step of the whole model for the batch: 32 x [RMSNorm, q/k/v proj, RoPE, KIVI decode attention +
cache update (two libkivi_b200 launches: q.K^T, p.V), o_proj, MLP], final norm, lm_head, logits all-gather (N > 1),
greedy argmax, cache advance.  The cache is pre-filled with synthetic K/V by the real prefill pack
kernels so that the K timed steps END at seq = 4096 tokens; weights are random-init fp16 (no
checkpoints offline).  N > 1: data-parallel replicas, batch 32 per GPU (weak scaling), one NCCL
all-gather of the logits per step.

One JSON line on stdout (rank 0).  `value` = whole-job tokens/s with inputs resident in HBM;
`e2e` = same metric through the public API with HOST buffers (token ids pinned -> H2D, logits D2H
every step); `roofline` = the decode-attention call (the dominant kernels of the hot path) against
the measured HBM peak; `cpu_baseline` = the reference's CPU fake-quant attention (oracle port of
models/utils_quant.py) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decode tokens/sec @ Llama-2-7B bs32 seq4k K2V2"
UNIT = "tokens/s"


# --------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: NVML polled from a thread every 5 ms (a 16-step
    timed region lasts ~0.1 s, too short for `nvidia-smi -lms`); falls back to one nvidia-smi query."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index: int = 0):
        self.gpu_index, self.sm, self.mask, self.power = gpu_index, [], 0, []
        self.h, self.nv, self.stop_flag, self.th, self.mx = None, None, False, None, None

    def _handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        self.nv = pynvml
        try:
            uuid = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu_index])
                except Exception:
                    pass
            return pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            self.h = self._handle()
            self.mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.h = None

    def stop(self):
        if self.h is not None:
            self.stop_flag = True
            self.th.join(timeout=1)
            if self.sm:
                return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.mx,
                        "reasons": sorted(n for n, bit in self.REASONS if self.mask & bit),
                        "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm), "source": "nvml 5 ms poll"}
        try:    # fallback: one nvidia-smi query right after the timed region
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                    "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU fake-quant attention on host cores
# --------------------------------------------------------------------------------------------------
def cpu_fake_quant_sample(batch: int, heads: int, kv_heads: int, T: int, g: int, bits: int, reps: int, layers: int):
    """One attention layer of fake-quant decode (oracle/fake_quant.py: models/utils_quant.py:167-217, :418-432,
    :498-563 restated) for `batch` sequences at kv length T; returns (tokens/s extrapolated to `layers`
    layers, seconds per layer-call, cores)."""
    import torch
    from oracle import fake_quant
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(0)
    Tq = T - T % g
    q = torch.randn((batch, heads, 1, 128), generator=gen, dtype=torch.float32)
    k = torch.randn((batch, kv_heads, Tq, 128), generator=gen, dtype=torch.float32)
    v = torch.randn((batch, kv_heads, Tq, 128), generator=gen, dtype=torch.float32)
    fake_quant.fake_quant_decode_attention(q, k, v, g, bits, bits)           # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fake_quant.fake_quant_decode_attention(q, k, v, g, bits, bits)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return batch / (t * layers), t, cores, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cfg = workload_config(args)
    b_sample = 2
    tps_list = []
    t_start = time.perf_counter()
    for _ in range(args.warmup):
        cpu_fake_quant_sample(b_sample, 32, 32, args.seq, 32, 2, 1, 32)
    for _ in range(args.steps):
        tps, t_layer, cores, threads = cpu_fake_quant_sample(b_sample, 32, 32, args.seq, 32, 2, 1, 32)
        tps_list.append(tps)
    value = statistics.median(tps_list)
    sample = (f"one attention layer of the reference's CPU fake-quant decode (models/utils_quant.py simulate paths, "
              f"oracle port) at kv length {args.seq}, {b_sample} sequences x 32 heads, fp32; tokens/s = {b_sample} / "
              f"(t_layer x 32 layers); linears excluded (attention hot path only)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * b_sample / value if value else None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_start}
    print(json.dumps(line))
    return 0


def workload_config(args):
    return {"workload": f"Llama-2-7B K2V2 g32 residual128, bs{args.batch} per GPU, decode steps ending at seq {args.seq} "
                        f"(cache pre-filled by the prefill pack kernels), 1xB200 per rank",
            "batch_per_gpu": args.batch, "seq_len": args.seq, "k_bits": 2, "v_bits": 2, "group_size": 32,
            "residual_length": 128, "parallelism": f"dp{args.gpus}",
            "l2": "per-step working set (13.5 GB weights + ~15 GB KV cache) >> 126 MB L2: inputs larger than L2"}


# --------------------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    os.environ.setdefault("NCCL_DEBUG", "WARN")              # keep NCCL's version banner off stdout: one JSON line only
    import torch
    from kivi_b200 import _lib, dist as kdist
    from kivi_b200.llama_kivi import LlamaForCausalLM_KIVI, default_config

    rank, ws, local = kdist.init()
    assert ws == args.gpus or ws == 1, f"--gpus {args.gpus} but WORLD_SIZE={ws}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, K, W, seq = args.batch, args.steps, max(args.warmup, 3), args.seq
    Bg = B * ws

    cfg = default_config(args.model)
    torch.manual_seed(0)
    with torch.device(dev):
        model = LlamaForCausalLM_KIVI(cfg).half()
    for p_ in model.parameters():
        p_.requires_grad_(False)
    model.eval()
    n_e2e = K
    total_steps = W + K + W + n_e2e + 4
    n0 = seq - (W + K)                                       # the K timed steps end at kv length `seq`
    model.init_cache(B, max_tokens=seq + total_steps + 8)
    model.prefill_synthetic(n0, seed=rank)
    cache = model.cache
    vocab = cfg.vocab_size

    ids = torch.randint(0, vocab, (B, 1), device=dev)
    lo, hi = kdist.shard_range(Bg, rank, ws)

    def step(tok):
        logits = model.decode_step(tok)                      # CUDA-graph replay of the whole step
        _, mine = kdist.greedy_next_tokens(logits, rank, ws, Bg)   # NCCL all-gather of the logits (N > 1) + argmax
        return mine.view(B, 1)

    # ---- warm-up (captures the graph), then the timed region
    for _ in range(W):
        ids = step(ids)
    torch.cuda.synchronize()
    # our launches per captured step = libkivi_b200 launches enqueued while capturing (the graph replays them)
    launches_per_step = getattr(model, "launches_per_step", None) or (2 * cfg.num_hidden_layers + 1)
    if os.environ.get("KIVI_PROFILE_STEPS"):                 # ncu --profile-from-start off: profile N steps, exit
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(int(os.environ["KIVI_PROFILE_STEPS"])):
            ids = step(ids)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0
    sampler = ClockSampler(local)
    kdist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        ids = step(ids)
    e1.record()
    torch.cuda.synchronize()
    kdist.barrier()
    ms = kdist.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    state_at_end = [cache.tk, cache.r, cache.tv, cache.L, cache.kv_len]
    value = Bg * K / (ms / 1e3)

    # ---- e2e: host token ids (pinned) -> H2D, step, logits D2H (pinned), every step
    ids_host = torch.empty((B, 1), dtype=torch.long).pin_memory()
    logits_host = torch.empty((B, vocab), dtype=torch.float32).pin_memory()
    ids_host.copy_(ids.cpu())
    ids_dev = torch.empty((B, 1), dtype=torch.long, device=dev)

    def step_e2e():
        ids_dev.copy_(ids_host, non_blocking=True)              # H2D: this step's token ids (pinned)
        logits = model.decode_step(ids_dev)
        toks, mine = kdist.greedy_next_tokens(logits, rank, ws, Bg)
        logits_host.copy_(logits, non_blocking=True)           # D2H: the step's result (logits of this shard)
        ids_host.copy_(mine.view(B, 1), non_blocking=True)     # D2H: sampled ids, fed back from the host next step
        torch.cuda.current_stream().synchronize()

    for _ in range(W):
        step_e2e()
    kdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n_e2e):
        step_e2e()
    e1.record()
    torch.cuda.synchronize()
    kdist.barrier()
    ms_e2e = kdist.max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    e2e_value = Bg * n_e2e / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: fused decode attention, timed alone with CUDA events on its stream
    roof = None
    if rank == 0:
        q = torch.randn((B, cfg.num_attention_heads, 128), device=dev, dtype=torch.float16)
        kn = torch.randn((B, cfg.num_key_value_heads, 128), device=dev, dtype=torch.float16)
        vn = torch.randn_like(kn)
        out = torch.empty_like(q)
        NL = cfg.num_hidden_layers
        while cache.r == cache.residual_length - 1:          # stay off the K-flush step (once per R steps)
            ids = step(ids)
        for l in range(NL):                                  # one cold pass over all layers (15 GB >> L2)
            cache.decode_attention(l, q, kn, vn, out=out)
        torch.cuda.synchronize()
        reps = 3
        e0.record()
        for _ in range(reps):
            for l in range(NL):
                cache.decode_attention(l, q, kn, vn, out=out)
        e1.record()
        torch.cuda.synchronize()
        per_launch_ms = e0.elapsed_time(e1) / (reps * NL)
        per_tok = 128 * (2 / 8 + 4 / 32)
        U = B * cfg.num_key_value_heads
        alg_bytes = U * ((cache.tk + cache.tv) * per_tok + (cache.r + cache.L) * 256) + \
            (2 * B * cfg.num_attention_heads + 2 * U) * 256
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_attention_ncu.json")) as f:      # ncu --set full, both kernels of the call
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        roof = {"kernel": "kivi_decode_attention_f16 = kivi::qk_kernel<2,1,32> + kivi::sv_kernel<2,2,1,32> (q.Kq^T + window + softmax "
                          "statistics | normalise + p.Vq + window + output + cache update), timed as one call",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "launch_ms": per_launch_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "state": [cache.tk, cache.r, cache.tv, cache.L],
                "share_of_step": per_launch_ms * NL / (ms / K)}

    # ---- cpu baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        tps, t_layer, cores, threads = cpu_fake_quant_sample(2, 32, 32, seq, 32, 2, 3, 32)
        cpu = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"oracle port of the reference's CPU fake-quant decode attention (models/utils_quant.py): one layer, "
                          f"2 sequences x 32 heads, kv length {seq}, fp32, {threads} torch threads, median of 3 "
                          f"({t_layer:.2f} s per layer-call); tokens/s = 2 / (t_layer x 32 layers), linears excluded")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 (fp32 accumulate; 2-bit codes)", "data": "synthetic", "config": workload_config(args),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": B * vocab * 4 + B * 8,
                        "ms_per_step": ms_e2e / n_e2e},
                "gpu_launches": launches_per_step * K,
                "gpu_launches_note": f"{launches_per_step} libkivi_b200 launches per step, counted by the library while the step was "
                                     f"captured and replayed from a CUDA graph: per layer q.K^T + p.V attention kernels, "
                                     f"add+RMSNorm x2, RoPE+split, SiLU*mul; final norm; cache advance (cuBLAS GEMMs not counted)",
                "roofline": roof, "cpu_baseline": cpu, "cache_state_after_timed": state_at_end,
                "note": "the timed steps end at seq 4096 and therefore include the once-per-128-steps K flush step (tk 3968 -> 4096); "
                        "the e2e steps follow at seq 4097..",
                "model": args.model, "global_batch": Bg}
        print(json.dumps(line))
    if ws > 1:
        import torch.distributed as td
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-2-7b")
    ap.add_argument("--batch", type=int, default=32, help="sequences per GPU")
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
