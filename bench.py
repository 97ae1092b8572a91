#!/usr/bin/env python
"""bench.py — sentences/s of the ReaLiSe multimodal hot path, forward + backward + optimizer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload = the configuration BASELINE.json's metric ("sentences/sec fwd+bwd seq_len=128") is quoted
on: configs[2] at N=1 (one optimizer step = train-mode forward with dropout 0.1 and batch-statistics BatchNorm,
backward, clip_grad_norm_(1.0), AdamW; batch 128 x seq_len 128; all three encoders + gate + output block +
21128-way classifier; bf16 operands / fp32 accumulate, fp32 master weights and moments) and configs[3] at N>1
(128 sentences per GPU, ONE NCCL all-reduce of the flat fp32 gradient buffer per step, weak scaling).
Synthetic batch (realise_b200.synth) and random-init weights of the full architecture.
`value` = device-resident inputs, CUDA events around the K steps, max over ranks.  `e2e` = the reference's
training-loop body (src/run.py:186-212) through the public API with a pinned HOST batch: H2D of the batch,
model(batch) -> loss.backward() -> optimizer.step() -> loss.item() (D2H).  The forward-only configs[1] number
(B=64, eval, CUDA graph) is reported as the secondary `forward_only` object.
`--impl reference` / `cpu_baseline` time the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py:
its SpellBertPho2ResArch3 and the loop body of src/run.py:186-212 with the vendored AdamW) on the host cores on a
bounded sample (`kind: "reference"`; the oracle port is the fallback when baseline/_ref is absent).  `incumbent` = the
same reference modules under torch eager on the B200 (fp32 and autocast-bf16), same batch.  One JSON line (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_TRAIN, B_FWD, SEQ_LEN = 128, 64, 128
METRIC = "sentences/sec fwd+bwd seq_len=128 (BASELINE configs[2]/[3]: train step fwd+bwd+clip+AdamW, B=128/GPU)"
FLOP_PER_SENTENCE_FWD = 59.39e9      # SURVEY.md §8(d), nominal, L=128
FLOP_PER_SENTENCE_TRAIN = 178.2e9    # 3 x forward


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic():
    for name in ("r02_train_traffic.json", "r01_train_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["gemm_dram_bytes_per_launch"])
        except (OSError, KeyError, ValueError):
            continue
    return None


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


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference train step (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------
class CpuTrainStep:
    """fwd (train mode, dropout 0.1, batch-stat BN) + autograd bwd + clip_grad_norm_(1.0) + AdamW with the
    vendored optimizer's semantics (transformers/optimization.py:113-169) on the oracle's state_dict."""

    def __init__(self, batch_sentences, seq_len):
        from oracle import realise_oracle as O
        from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
        O.FAST = True  # ATen fused CPU kernels, like the reference's nn.Modules
        self.O, self.cfg = O, ArchConfig()
        sd = synth_state_dict(self.cfg, seed=0)
        sd["classifier.weight"] = sd["bert.embeddings.word_embeddings.weight"]   # tie_cls_weight (src/models.py:700)
        frozen = ("char_images", "num_batches_tracked", "running_")
        self.leaves = []
        for k, v in sd.items():
            if v.dtype.is_floating_point and not any(f in k for f in frozen) and k != "classifier.weight":
                v.requires_grad_(True)
                self.leaves.append(v)
        self.sd = sd
        self.m = [torch.zeros_like(p) for p in self.leaves]
        self.v = [torch.zeros_like(p) for p in self.leaves]
        self.t = 0
        self.B = batch_sentences
        self.batch = synth_batch(batch_sentences, seq_len, seed=1, ragged=False, with_labels=True)

    def __call__(self):
        loss = self.O.forward(self.sd, self.batch, self.cfg, train=True)[0]
        loss.backward()
        ps = [p for p in self.leaves if p.grad is not None]
        gs = [p.grad for p in ps]
        with torch.no_grad():
            torch.nn.utils.clip_grad_norm_(ps, 1.0)
            self.t += 1
            ms = [m for m, p in zip(self.m, self.leaves) if p.grad is not None]
            vs = [v for v, p in zip(self.v, self.leaves) if p.grad is not None]
            torch._foreach_mul_(ms, 0.9)
            torch._foreach_add_(ms, gs, alpha=0.1)
            torch._foreach_mul_(vs, 0.999)
            torch._foreach_addcmul_(vs, gs, gs, value=0.001)
            step = 5e-5 * math.sqrt(1.0 - 0.999 ** self.t) / (1.0 - 0.9 ** self.t)
            den = torch._foreach_sqrt(vs)
            torch._foreach_add_(den, 1e-8)
            torch._foreach_addcdiv_(ps, ms, den, value=-step)
            for p in ps:
                p.grad = None
        return float(loss.detach())


def cpu_train_rate(budget_s, max_steps, sample_b=4, warmup=1):
    """sentences/s of the CPU oracle PORT's train step on a bounded sample (fallback when baseline/_ref is absent);
    returns (rate, steps, threads, sample text, seconds)."""
    maxt = host_threads()
    torch.set_num_threads(min(maxt, 16))
    small = CpuTrainStep(1, SEQ_LEN)
    small()                                                        # page in, warm the allocator
    pick_threads(small, sorted({min(maxt, c) for c in (8, 16, 32, 64)}))
    del small
    st = CpuTrainStep(sample_b, SEQ_LEN)
    for _ in range(warmup):
        st()
    t0, n = time.perf_counter(), 0
    while n < max_steps:
        st()
        n += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    sample = (f"{n} x train step (fwd+bwd+clip+AdamW, dropout 0.1) of {sample_b} sentences x {SEQ_LEN} tokens, "
              f"oracle/realise_oracle.py FAST mode (ATen CPU kernels), fp32")
    return n * sample_b / dt, n, torch.get_num_threads(), sample, dt


def reference_cpu_rate(steps, warmup, budget_s):
    """sentences/s of the UNMODIFIED reference (baseline/_ref: its SpellBertPho2ResArch3 + the loop body of
    src/run.py:186-212 with clip_grad_norm_ and the vendored AdamW) on the host cores.  Each step is a bounded sample of
    the BASELINE configs[2] workload: the largest batch in {4 .. 128} sentences x 128 tokens for which `warmup + steps`
    steps fit the time budget.  Returns (rate, threads, sample text, seconds, sample batch)."""
    from baseline import ref_model
    from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
    maxt = host_threads()
    sd = synth_state_dict(ArchConfig(), seed=0)
    model = ref_model.build(sd)
    probe = ref_model.TrainStep(model, synth_batch(2, SEQ_LEN, seed=1, ragged=False))
    torch.set_num_threads(min(maxt, 16))
    probe()                                                        # page in, warm the allocator
    pick_threads(probe, sorted({min(maxt, c) for c in (8, 16, 32, 64)}))
    t0 = time.perf_counter()
    probe()
    t2 = time.perf_counter() - t0
    probe.batch = synth_batch(4, SEQ_LEN, seed=1, ragged=False)
    t0 = time.perf_counter()
    probe()
    t4 = time.perf_counter() - t0
    per_sent = max((t4 - t2) / 2.0, 1e-3)
    fixed = max(t2 - 2 * per_sent, 0.0)                            # optimizer + clip: independent of the batch
    sample_b = 4
    for b in (8, 16, 32, 64, 128):
        if (steps + warmup) * (fixed + b * per_sent) <= budget_s:
            sample_b = b
    st = ref_model.TrainStep(model, synth_batch(sample_b, SEQ_LEN, seed=1, ragged=False))
    st.opt = probe.opt
    for _ in range(warmup):
        st()
    t0 = time.perf_counter()
    for _ in range(steps):
        st()
    dt = time.perf_counter() - t0
    sample = (f"{steps} x the reference's own train step (SpellBertPho2ResArch3.forward + backward + clip_grad_norm_ + vendored "
              f"AdamW, dropout 0.1, fp32; baseline/_ref) on {sample_b} sentences x {SEQ_LEN} tokens after {warmup} warm-up steps")
    return steps * sample_b / dt, torch.get_num_threads(), sample, dt, sample_b


def run_reference(args, rank, world):
    if rank != 0:
        return
    from baseline import ref_model
    if ref_model.available():
        steps, warm = args.steps, args.warmup
        value, threads, sample, dt, sample_b = reference_cpu_rate(steps, warm, budget_s=args.ref_budget)
        kind = "reference"
    else:
        value, steps, threads, sample, dt = cpu_train_rate(budget_s=150.0, max_steps=args.steps, warmup=1)
        kind, warm, sample_b = "port", 1, 4
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sentences/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: train step fwd+bwd+clip+AdamW of the full SpellBertPho2ResArch3, "
                               f"seq_len 128; each step = bounded sample of {sample_b} sentences on the host CPU",
                   "global_batch": sample_b, "seq_len": SEQ_LEN},
        "cpu_baseline": {"value": value, "unit": "sentences/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def measure_incumbent(dev, steps=3):
    """The comparator that means something for a GPU path: the UNCHANGED reference modules (baseline/_ref) under stock
    torch eager on the same B200, same batch (B=128 x L=128), the reference's own loop body — in fp32 (its published
    recipe) and under torch.autocast(bfloat16).  CUDA events, 1 warm-up + `steps` timed steps each."""
    from baseline import ref_model
    from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
    if not ref_model.available():
        return {"unavailable": "baseline/_ref not installed (baseline/install_ref.py)"}
    out = {"workload": f"reference SpellBertPho2ResArch3 on the B200, torch {torch.__version__} eager, train step B={B_TRAIN} x "
                       f"L={SEQ_LEN} (fwd+bwd+clip_grad_norm_+vendored AdamW), 1 warm-up + {steps} timed steps, CUDA events"}
    host = synth_batch(B_TRAIN, SEQ_LEN, seed=4321, ragged=False, with_labels=True)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    for name, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        try:
            model = ref_model.build(synth_state_dict(ArchConfig(), seed=0), device=dev)
            st = ref_model.TrainStep(model, batch, autocast_dtype=dt)
            st()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = st()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": B_TRAIN / (ms * 1e-3), "unit": "sentences/s", "ms_per_step": ms, "loss": float(loss.detach())}
        except Exception as e:  # noqa: BLE001 — a comparator must never take the headline down
            out[name] = {"error": repr(e)[:300]}
        finally:
            model = st = None
            torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def _agg(prof):
    agg = {}
    for kind, work, e0, e1, *_ in prof:
        a = agg.setdefault(kind, [0.0, 0.0, 0])
        a[0] += work
        a[1] += e0.elapsed_time(e1) * 1e-3
        a[2] += 1
    return agg


def build_train(dev, rank, world):
    from realise_b200.ddp import DataParallel
    from realise_b200.model import SpellBertPho2ResArch3Abla
    from realise_b200.optim import FusedAdamW
    from realise_b200.synth import ArchConfig
    cfg = ArchConfig(with_pho="yes", with_res="yes")       # == SpellBertPho2ResArch3 (src/models.py:652)
    torch.manual_seed(0)
    model = SpellBertPho2ResArch3Abla(cfg)
    model.tie_cls_weight()
    model.train().to(dev)
    if world > 1:
        dp = DataParallel(model)
        dp.broadcast_parameters()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    nd = [p for n, p in named if "bias" in n or "LayerNorm.weight" in n]                 # src/run.py:146-151
    dc = [p for n, p in named if not ("bias" in n or "LayerNorm.weight" in n)]
    opt = FusedAdamW([{"params": dc, "weight_decay": 0.0}, {"params": nd, "weight_decay": 0.0}], lr=5e-5, eps=1e-8,
                     max_grad_norm=1.0, model=model)
    return model, opt, sum(p.numel() for _, p in named)


def measure_train(args, dev, rank, world, dist, peaks):
    from realise_b200 import ops
    from realise_b200.synth import synth_batch
    B, L = B_TRAIN, SEQ_LEN
    model, opt, n_params = build_train(dev, rank, world)
    host = synth_batch(B, L, seed=4321 + rank, ragged=False, with_labels=True)
    db = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    db["pho_lens"] = torch.tensor(host["pho_lens"], dtype=torch.int32, device=dev)

    def eager_step(b=db):
        loss = model(b)[0]
        loss.backward()
        opt.step()
        return loss

    # the repo's public training-step API: fwd + bwd + (all-reduce) + clip + AdamW as one CUDA-graph replay
    from realise_b200.graphed import GraphedTrainStep
    gstep = GraphedTrainStep(model, opt)
    graphed = not args.eager

    def step(b=db):
        return gstep(b) if graphed else eager_step(b)

    W = max(args.warmup, 3)
    eager_step()
    torch.cuda.synchronize()
    # ---- launches per step + per-kernel roofline pass (CUDA events around every C-ABI call, eager) ----
    n0 = ops.LAUNCHES
    eager_step()
    launches_per_step = ops.LAUNCHES - n0 + 2          # + the two optimizer kernels (sum of squares, clip+AdamW)
    ops._prof = []
    eager_step()
    torch.cuda.synchronize()
    prof, ops._prof = ops._prof, None
    if args.dump_prof and rank == 0:
        with open(args.dump_prof, "w") as fh:
            json.dump([{"kind": k, "detail": list(d) if d else None, "us": round(e0.elapsed_time(e1) * 1e3, 2), "work": w}
                       for k, w, e0, e1, d in prof], fh)
    try:
        for _ in range(W):
            step()                                         # first call of the shape is eager, the second captures
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001 — capture unsupported in this environment: time the eager loop instead
        sys.stderr.write(f"bench: CUDA-graph capture of the train step failed ({e!r}); falling back to the eager loop\n")
        graphed = False
        torch.cuda.synchronize()
        for _ in range(W):
            step()
        torch.cuda.synchronize()
    agg = _agg(prof)
    gw = agg.get("gemm", [0, 0, 0])[0] + agg.get("conv_gemm", [0, 0, 0])[0]
    gt = agg.get("gemm", [0, 1e-9, 0])[1] + agg.get("conv_gemm", [0, 0, 0])[1]
    gn = agg.get("gemm", [0, 0, 0])[2] + agg.get("conv_gemm", [0, 0, 0])[2]
    gemm_tf = gw / gt / 1e12

    # ---- timed region: K optimizer steps, device-resident batch, CUDA events, max over ranks ----
    clocks = ClockSampler(dev.index)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clk = clocks.stop()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    value = world * B * args.steps / (total_ms * 1e-3)
    final_loss = float(loss.item())

    # ---- e2e: pinned host batch -> H2D -> forward/backward/optimizer through the public API -> loss.item() ----
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values() if torch.is_tensor(v)) + 4 * len(host["pho_lens"])

    def e2e_step():
        b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()}
        return step(b).item()                     # pho_lens stays a Python list (src/run.py:189); .item() = D2H + sync

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / t.item()
    peak_mem = torch.cuda.max_memory_allocated(dev) / 1e9
    att, attb = agg.get("attention"), agg.get("attention_bwd")
    # BertSelfAttention (modeling_bert.py:220-263) = fused QKV projection + attention core, forward and fwd+bwd: the
    # "attention-GEMM roofline" north_star asks for (3.932 MFLOP/token/layer forward)
    N_tok, H3, Hd = B * L, 3 * 768, 768
    qkv_f = [0.0, 0.0]
    qkv_b = [0.0, 0.0]
    for kind, work, e0_, e1_, detail in prof:
        if kind != "gemm" or not detail:
            continue
        m_, n_, k_ = detail[0], detail[1], detail[2]
        dt_ = e0_.elapsed_time(e1_) * 1e-3
        if (m_, n_, k_) == (N_tok, H3, Hd):
            qkv_f[0] += work; qkv_f[1] += dt_
        elif (m_, n_, k_) in ((N_tok, Hd, H3), (H3, Hd, N_tok)):
            qkv_b[0] += work; qkv_b[1] += dt_
    bsa = None
    if att and attb and qkv_f[1] > 0:
        f_w, f_t = qkv_f[0] + att[0], qkv_f[1] + att[1]
        a_w, a_t = f_w + qkv_b[0] + attb[0], f_t + qkv_b[1] + attb[1]
        bsa = {"forward_tflops": f_w / f_t / 1e12, "forward_frac_of_peak": f_w / f_t / 1e12 / peaks["tf_sustained"],
               "fwd_bwd_tflops": a_w / a_t / 1e12, "fwd_bwd_frac_of_peak": a_w / a_t / 1e12 / peaks["tf_sustained"],
               "ms_per_step": a_t * 1e3,
               "what": "fused QKV GEMM + attention core (19 layers), executed FLOPs over CUDA-event time, against the measured "
                       "sustained bf16 peak; ncu tensor-pipe-active per kernel: profiles/r02_ncu_full_kernels.json"}
    out = {
        "value": value, "ms_per_step": total_ms / args.steps, "warmup": W, "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "sentences/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches_per_step * args.steps,
        "model_tflops": value * FLOP_PER_SENTENCE_TRAIN / 1e12 / world,
        "roofline": {"bound": "tensor", "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tf / peaks["tf_sustained"],
                     # dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, from the committed ncu pass over one
                     # train step (profiles/r02_train_traffic.json); null when that file is absent
                     "traffic": load_traffic(),
                     "kernel": "gemm_bf16_kernel / gemm2_bf16_kernel / conv64_halo_kernel (tcgen05 GEMM, implicit-GEMM conv, split-K wgrad): "
                               f"executed 2*M*N*K over CUDA-event time of its {gn} launches in one train step",
                     "peak_source": peaks["source"] + " bf16 sustained"},
        "roofline_detail": {
            "gemm_ms_per_step": gt * 1e3,
            "attention_fwd_tflops": None if not att else att[0] / att[1] / 1e12,
            "attention_bwd_tflops": None if not attb else attb[0] / attb[1] / 1e12,
            "bert_self_attention": bsa,
            "per_kernel_ms_per_step": {k: round(v[1] * 1e3, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
            "timed_kernels_ms": round(sum(v[1] for v in agg.values()) * 1e3, 3),
        },
        "config": {"workload": "BASELINE configs[2] (N=1) / configs[3] (N>1): one optimizer step of the full "
                               "SpellBertPho2ResArch3 (BERT 12L + pinyin GRU/4L + glyph CharResNet + gate + 3L output "
                               "block + tied classifier): train-mode fwd (dropout 0.1, batch-stat BN) + bwd + clip(1.0) "
                               "+ AdamW; random-init full-size weights",
                   "global_batch": world * B, "seq_len": L, "parallelism": f"dp{world} (batch-sharded, "
                   + ("one NCCL all-reduce of the flat fp32 gradient buffer per step)" if world > 1 else "no collective)"),
                   "trainable_params": n_params,
                   "l2": f"no explicit flush: a step streams {peak_mem:.1f} GB of activations/gradients/optimizer state "
                         "(>> 126 MB L2) between any two uses of the same data",
                   "cuda_graph": graphed},
        "final_loss": final_loss, "peak_mem_gb": peak_mem,
    }
    del model, opt
    torch.cuda.empty_cache()
    return out


def measure_forward(args, dev, rank, world, dist, peaks):
    """Secondary object: BASELINE configs[1], forward-only eval, B=64 x L=128 per GPU, CUDA graph, L2 flushed."""
    from realise_b200 import ops
    from realise_b200.model import SpellBertPho2ResArch3
    from realise_b200.synth import ArchConfig, synth_batch, synth_state_dict
    cfg = ArchConfig()
    model = SpellBertPho2ResArch3(cfg)
    model.tie_cls_weight()
    model.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
    model.eval().to(dev)
    B, L = B_FWD, SEQ_LEN
    host = synth_batch(B, L, seed=1234 + rank, ragged=False, with_labels=True)
    dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    dbatch["pho_lens"] = torch.tensor(host["pho_lens"], dtype=torch.int32, device=dev)
    fwd_batch = {k: v for k, v in dbatch.items() if k not in ("tgt_idx", "loss_masks")}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    steps = min(args.steps, 20)
    with torch.no_grad():
        model.use_cuda_graph = False
        model(fwd_batch)
        ops._prof = []
        model(fwd_batch)
        torch.cuda.synchronize()
        prof, ops._prof = ops._prof, None
        model.use_cuda_graph = True
        for _ in range(3):
            model(fwd_batch)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(fwd_batch)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = world * B * steps / (t.item() * 1e-3)
    agg = _agg(prof)
    gw = agg.get("gemm", [0, 0, 0])[0] + agg.get("conv_gemm", [0, 0, 0])[0]
    gt = agg.get("gemm", [0, 1e-9, 0])[1] + agg.get("conv_gemm", [0, 0, 0])[1]
    att = agg.get("attention")
    stem = agg.get("glyph_block1") or agg.get("glyph_stem")
    out = {"value": value, "unit": "sentences/s", "ms_per_step": t.item() / steps, "steps": steps,
           "workload": "BASELINE configs[1]: forward-only eval, batch 64/GPU x seq_len 128, CUDA graph, 256 MB L2 flush "
                       "between steps", "model_tflops": value * FLOP_PER_SENTENCE_FWD / 1e12 / world,
           "gemm_tflops": gw / gt / 1e12, "gemm_frac_of_peak": gw / gt / 1e12 / peaks["tf_sustained"],
           "attention_tflops": None if not att else att[0] / att[1] / 1e12,
           "glyph_block1_gbs": None if not stem else stem[0] / stem[1] / 1e9,
           "glyph_block1_frac_of_hbm_peak": None if not stem else stem[0] / stem[1] / 1e9 / peaks["hbm_gbs"],
           "per_kernel_ms_per_step": {k: round(v[1] * 1e3, 3) for k, v in agg.items()}}
    del model
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    tr = measure_train(args, dev, rank, world, dist, peaks)
    fwd = None
    if not args.no_forward:
        try:
            fwd = measure_forward(args, dev, rank, world, dist, peaks)
        except Exception as e:  # noqa: BLE001 — the secondary object must never take the headline down
            fwd = {"error": repr(e)[:300]}
    inc = None
    if world == 1 and not args.no_incumbent:
        inc = measure_incumbent(dev)
        for k in ("fp32", "bf16_autocast"):
            if isinstance(inc.get(k), dict) and "value" in inc[k]:
                inc[k]["ours_over_incumbent"] = tr["value"] / inc[k]["value"]
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from baseline import ref_model
        if ref_model.available():
            rate, threads, sample, _, _ = reference_cpu_rate(steps=3, warmup=1, budget_s=25.0)
            cpu = {"value": rate, "unit": "sentences/s", "cores": threads, "kind": "reference", "sample": sample}
        else:
            rate, reps, threads, sample, _ = cpu_train_rate(budget_s=20.0, max_steps=20, warmup=1)
            cpu = {"value": rate, "unit": "sentences/s", "cores": threads, "kind": "port", "sample": sample}
    line = {
        "metric": METRIC, "value": tr["value"], "unit": "sentences/s", "n_gpus": world, "steps": args.steps,
        "warmup": tr["warmup"], "ms_per_step": tr["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": tr["config"], "model_tflops_per_gpu": tr["model_tflops"],
        "roofline": tr["roofline"], "roofline_detail": tr["roofline_detail"],
        "cpu_baseline": cpu, "e2e": tr["e2e"], "gpu_launches": tr["gpu_launches"], "clocks": tr["clocks"],
        "final_loss": tr["final_loss"], "peak_mem_gb": tr["peak_mem_gb"], "forward_only": fwd, "incumbent": inc,
    }
    emit(line)


def main():
    # the contract is ONE JSON line on stdout: anything a library prints there (NCCL's version banner, warnings) is
    # sent to stderr instead; the line itself goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time the eager Python loop instead of the CUDA-graph step")
    ap.add_argument("--no-forward", action="store_true", help="skip the secondary forward-only (configs[1]) measurement")
    ap.add_argument("--ref-budget", type=float, default=170.0,
                    help="--impl reference: seconds the warm-up + timed CPU steps may take (sizes the bounded sample)")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the reference-on-the-B200 (torch eager) comparator")
    ap.add_argument("--dump-prof", default=None, help="write the per-call timeline of the eager profiling pass (kind, shape key, "
                    "us, executed work) of one train step to this JSON file")
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
