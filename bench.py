#!/usr/bin/env python
"""bench.py — sentences/s of the ReaLiSe multimodal hot path (SpellBertPho2ResArch3.forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 1 GPU = forward-only eval step, batch 64 x seq_len 128, all
three encoders + fusion + classifier, bf16 operands / fp32 accumulate, synthetic batch
(realise_b200.synth) and seeded random weights of the full architecture (21128 vocab, 12+4+3
transformer layers, 3-font glyph table).  N GPUs: one process per GPU, each its own batch of 64
(sentences are independent: no data-path collective), whole-job value = N*64*K / max-rank time.
One JSON line on stdout (rank 0).  `--impl reference` times the CPU oracle port of the reference
on the host cores for the same metric/config on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, SEQ_LEN = 64, 128
METRIC = "sentences/sec fwd seq_len=128 (BASELINE configs[1]: forward-only B=64 L=128, all encoders + fusion)"
FLOP_PER_SENTENCE_FWD = 59.39e9  # SURVEY.md §8(d), nominal, L=128


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed to run on (affinity / cgroup
    quota), not the machine's core count — oversubscribing a shared host makes the baseline slower."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def pick_threads(fn, candidates):
    """Calibrate: run fn() once per candidate thread count and keep the fastest (best for the CPU arm)."""
    best, best_t = candidates[0], float("inf")
    for n in candidates:
        torch.set_num_threads(n)
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def cpu_oracle_rate(batch_sentences, seq_len, budget_s, threads=None):
    """sentences/s of the CPU oracle (port of the reference forward) on a bounded sample."""
    from oracle import realise_oracle as O
    from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
    O.FAST = True  # ATen fused CPU kernels, like the reference's nn.Modules
    if threads:
        torch.set_num_threads(threads)
    cfg = ArchConfig()
    sd = synth_state_dict(cfg, seed=0)
    batch = synth_batch(batch_sentences, seq_len, seed=1, ragged=False, with_labels=False)
    with torch.no_grad():
        torch.set_num_threads(min(threads or host_threads(), 16))
        O.forward(sd, synth_batch(2, seq_len, seed=2, ragged=False, with_labels=False), cfg)  # warm-up
        maxt = threads or host_threads()
        small = synth_batch(2, seq_len, seed=2, ragged=False, with_labels=False)  # cheap calibration batch
        pick_threads(lambda: O.forward(sd, small, cfg), sorted({min(maxt, c) for c in (8, 16, 32, 64)}))
        t0, n = time.perf_counter(), 0
        while True:
            O.forward(sd, batch, cfg)
            n += 1
            if time.perf_counter() - t0 >= budget_s or n >= 50:
                break
        dt = time.perf_counter() - t0
    return n * batch_sentences / dt, n, sd, cfg


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import realise_oracle as O
    from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
    O.FAST = True  # ATen fused CPU kernels, like the reference's nn.Modules
    cfg = ArchConfig()
    sd = synth_state_dict(cfg, seed=0)
    sample_b = 8  # bounded sample of the 64-sentence step
    batch = synth_batch(sample_b, SEQ_LEN, seed=1, ragged=False, with_labels=False)
    with torch.no_grad():
        maxt = host_threads()
        small = synth_batch(2, SEQ_LEN, seed=2, ragged=False, with_labels=False)  # cheap calibration batch
        torch.set_num_threads(min(maxt, 16))
        O.forward(sd, small, cfg)
        pick_threads(lambda: O.forward(sd, small, cfg), sorted({min(maxt, c) for c in (8, 16, 32, 64)}))
        for _ in range(max(0, min(args.warmup, 2) - 1)):
            O.forward(sd, batch, cfg)
        t0 = time.perf_counter()
        steps = 0
        for _ in range(args.steps):
            O.forward(sd, batch, cfg)
            steps += 1
            if time.perf_counter() - t0 > 150:
                break
        dt = time.perf_counter() - t0
    value = steps * sample_b / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sentences/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] forward-only B=64 L=128; each step = bounded sample of 8 sentences on CPU",
                   "global_batch": sample_b, "seq_len": SEQ_LEN},
        "cpu_baseline": {"value": value, "unit": "sentences/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} x forward of {sample_b} sentences x {SEQ_LEN} tokens, oracle/realise_oracle.py"},
        "e2e": {"value": value, "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_train(dev, rank, world, dist, steps):
    """Secondary line: one optimizer step (fwd + bwd + gradient all-reduce + clip + AdamW) of the full
    SpellBertPho2ResArch3 (all three encoders), batch 128/GPU x seq_len 128, dropout 0.1, batch-stat BatchNorm —
    BASELINE configs[2] (1 GPU) / configs[3] (N GPUs, weak scaling)."""
    from realise_b200.ddp import DataParallel
    from realise_b200.model import SpellBertPho2ResArch3Abla
    from realise_b200.optim import FusedAdamW
    from realise_b200.synth import ArchConfig, synth_batch
    B, L = 128, SEQ_LEN
    cfg = ArchConfig(with_pho="yes", with_res="yes")
    torch.manual_seed(0)
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.train().to(dev)
    if world > 1:
        dp = DataParallel(model)
        dp.broadcast_parameters()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    nd = [p for n, p in named if "bias" in n or "LayerNorm.weight" in n]
    dc = [p for n, p in named if not ("bias" in n or "LayerNorm.weight" in n)]
    opt = FusedAdamW([{"params": dc, "weight_decay": 0.0}, {"params": nd, "weight_decay": 0.0}], lr=5e-5, eps=1e-8,
                     max_grad_norm=1.0, model=model)
    host = synth_batch(B, L, seed=4321 + rank, ragged=False, with_labels=True)
    db = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    db["pho_lens"] = torch.tensor(host["pho_lens"], dtype=torch.int32, device=dev)

    def step():
        loss = model(db)[0]
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    out = {"value": world * B / ms * 1e3, "unit": "sentences/s", "ms_per_step": ms, "steps": steps,
           "config": {"workload": "BASELINE configs[2]: train step fwd+bwd+clip+AdamW of the full SpellBertPho2ResArch3 "
                                  "(BERT 12L + pinyin GRU/4L + glyph CharResNet + gate + 3L output block + classifier)",
                      "global_batch": world * B, "seq_len": L, "dropout": 0.1, "trainable_params": sum(p.numel() for _, p in named),
                      "grad_allreduce": "one NCCL all-reduce over the flat fp32 gradient buffer" if world > 1 else "none (1 GPU)"},
           "final_loss": float(loss.item()), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    del model, opt
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    from realise_b200 import ops
    from realise_b200.model import SpellBertPho2ResArch3
    from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    cfg = ArchConfig()
    sd = synth_state_dict(cfg, seed=0)
    model = SpellBertPho2ResArch3(cfg)
    model.tie_cls_weight()
    model.load_state_dict(sd, strict=True)
    model.eval().to(dev)
    B, L = B_PER_GPU, SEQ_LEN
    host = synth_batch(B, L, seed=1234 + rank, ragged=False, with_labels=True)
    dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    dbatch["pho_lens"] = torch.tensor(host["pho_lens"], dtype=torch.int32, device=dev)
    fwd_batch = {k: v for k, v in dbatch.items() if k not in ("tgt_idx", "loss_masks")}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- launches per step (eager pass) + per-kernel roofline pass (CUDA events per launch) ----
    with torch.no_grad():
        model.use_cuda_graph = False
        model(fwd_batch)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES
        model(fwd_batch)
        launches_per_step = ops.LAUNCHES - n0
        ops._prof = []
        for _ in range(3):
            model(fwd_batch)
        torch.cuda.synchronize()
        prof, ops._prof = ops._prof, None
        model.use_cuda_graph = True
    agg = {}
    for kind, work, e0, e1 in prof:
        a = agg.setdefault(kind, [0.0, 0.0, 0])
        a[0] += work
        a[1] += e0.elapsed_time(e1) * 1e-3
        a[2] += 1
    gemm_w = agg.get("gemm", [0, 0, 0])[0] + agg.get("conv_gemm", [0, 0, 0])[0]
    gemm_t = agg.get("gemm", [0, 1e-9, 0])[1] + agg.get("conv_gemm", [0, 0, 0])[1]
    gemm_n = agg.get("gemm", [0, 0, 0])[2] + agg.get("conv_gemm", [0, 0, 0])[2]
    gemm_tf = gemm_w / gemm_t / 1e12
    detail = {k: {"work": v[0] / 3, "ms": v[1] / 3 * 1e3, "launches": v[2] // 3} for k, v in agg.items()}
    att = agg.get("attention")
    stem = agg.get("glyph_block1") or agg.get("glyph_stem")

    # ---- timed region: K steps, device-resident inputs, CUDA events, L2 flushed between steps ----
    clocks = ClockSampler(local_rank)
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            model(fwd_batch)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        clocks.start()
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(fwd_batch)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        clk = clocks.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers in, predictions + loss out, through the public forward(batch) API ----
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    preds = torch.empty(B * L, dtype=torch.int64, device=dev)
    preds_host = torch.empty(B * L, dtype=torch.int64).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in pinned.values() if torch.is_tensor(v)) + 4 * len(host["pho_lens"])
    d2h = preds_host.numel() * 8 + 4

    def e2e_step():
        b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()}
        loss, logits = model(b)                       # pho_lens stays a Python list, like src/run.py:189
        ops.argmax_rows(logits.view(B * L, -1), preds)
        preds_host.copy_(preds, non_blocking=True)
        return loss.item()                            # D2H + sync, like src/run.py:202

    with torch.no_grad():
        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / t.item()

    train = None
    if not args.no_train:
        try:
            del model
            torch.cuda.empty_cache()
            train = measure_train(dev, rank, world, dist, steps=max(3, min(args.steps, 10)))
        except Exception as e:  # noqa: BLE001 — the secondary line must never take the headline down
            train = {"error": repr(e)[:300]}
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, reps, _, _ = cpu_oracle_rate(8, L, budget_s=15.0)
        cpu = {"value": rate, "unit": "sentences/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{reps} x eval forward of 8 sentences x {L} tokens (oracle/realise_oracle.py, fp32, all host threads)"}
    line = {
        "metric": METRIC, "value": value, "unit": "sentences/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: forward-only, batch 64/GPU x seq_len 128, 3 encoders + gate + "
                               "3-layer output block + 21128-way classifier; random-init full-size weights",
                   "global_batch": world * B, "seq_len": L, "parallelism": f"dp{world} (batch-sharded, no collective in forward)",
                   "l2": "256 MB buffer written between timed steps (L2 flush)", "cuda_graph": True},
        "model_tflops": value * FLOP_PER_SENTENCE_FWD / 1e12,
        "roofline": {"bound": "tensor", "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tf / peaks["tf_sustained"],
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 4 GEMM launches of one
                     # transformer layer in profiles/r01_ncu_full_summary.json (ncu --set full, round 1)
                     "traffic": 40.2e6,
                     "kernel": "gemm_bf16_kernel (tcgen05 GEMM + implicit-GEMM conv), executed 2*M*N*K over CUDA-event time, "
                               f"{gemm_n // 3} launches/step", "peak_source": peaks["source"] + " bf16 sustained"},
        "roofline_detail": {
            "attention_tensor": None if not att else {"achieved_tflops": att[0] / att[1] / 1e12,
                                                      "frac_of_peak": att[0] / att[1] / 1e12 / peaks["tf_sustained"]},
            "glyph_conv_hbm": None if not stem else {"kernel": "glyph_block1_kernel (gather + res_block1 fused)",
                                                     "achieved_gbs": stem[0] / stem[1] / 1e9,
                                                     "frac_of_peak": stem[0] / stem[1] / 1e9 / peaks["hbm_gbs"]},
            "per_kernel_ms_per_step": {k: round(v["ms"], 4) for k, v in detail.items()},
        },
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "sentences/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "train_step": train,
        "gpu_launches": (launches_per_step) * args.steps,
        "clocks": clk,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
