"""bench.py -- headline benchmark of the NeuSpeech hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): EEG train samples/sec, Whisper-base, eeg_ch=208, B=64 per GPU, L=32 labels, bf16 compute with fp32
master LoRA/stem weights, LoRA r=32 on the 36 encoder linears + 3 trainable stem convs, augmentation1 (identity values:
pad + cast + layout pass).  One "step" = augmentation pass + forward + loss + backward + (N>1: NCCL all-reduce of the flat
trainable gradient) + clip + AdamW -- the whole Trainer.training_step of the reference, nothing skipped.

value   : samples/s with the batch already resident in HBM (CUDA events, max over ranks)
e2e     : the same step through the public module API with HOST (pinned) inputs: H2D of the batch and D2H of the loss inside
          the timed region
roofline: the kernel family with the largest share of the step (tensor bound): algorithmic FLOPs / CUDA-event time of those
          launches inside a profiled step, against MEASURED_PEAKS.json (sustained bf16 figure: timed inside a long step)
cpu_baseline / --impl reference: the oracle (oracle/whisper_eeg.py = CPU restatement of the reference path; PEFT/accelerate
          are not installable here so the reference's own finetune.py cannot run) timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

FLOP_PER_SAMPLE = {208: 267.92e9, 273: 270.31e9}      # SURVEY.md 8(d) / BASELINE.md section 4 (Whisper-base, L=32, r=32)
METRIC = "EEG train samples/sec (Whisper-base, eeg_ch=208)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], 0.0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_host_batch(dims, B, L, seed):
    """Gwilliams-shaped synthetic batch on the host: 200 Hz signal in [-1,1] of random length, zero tail to 30 s."""
    g = torch.Generator().manual_seed(seed)
    x = (0.3 * torch.randn(B, dims.eeg_ch, dims.T, generator=g)).clamp_(-1, 1)
    for b in range(B):
        n = int(torch.randint(400, 5000, (1,), generator=g))
        x[b, :, n:] = 0
    labels = torch.randint(0, 50257, (B, L), generator=g)
    labels[:, -4:] = -100
    return x, labels


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle)
def cpu_train_steps(B, L, steps, warmup, eeg_ch=208, threads=None, lora_dropout=0.05):
    """Times oracle.train_step (fwd + loss + autograd bwd + clip + AdamW, LoRA r=32, LoRA-branch dropout) on the host.
    -> (samples/s, cores, s/step)"""
    from oracle import whisper_eeg as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    dims = O.Dims(eeg_ch=eeg_ch)
    P = O.init_params(dims, seed=0)
    lora = O.init_lora(dims, seed=1)
    st = O.AdamWState()
    x, labels = O.synthetic_batch(dims, B=B, L=L, seed=1)
    seed = 1

    def one():
        nonlocal seed
        if lora_dropout > 0:
            seed = O.next_dropout_seed(seed)
            lora["__dropout__"] = (lora_dropout, seed)
        O.train_step(x, labels, P, dims, lora, st, lr=1e-3)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return B * steps / dt, threads, dt / max(steps, 1)


def workload_name(eeg_ch, B, L, p=0.05):
    return (f"Gwilliams-shaped LoRA fine-tune step: Whisper-base, eeg_ch={eeg_ch}, B={B}/GPU, L={L}, "
            f"LoRA r=32 alpha=64 lora_dropout={p:g} on 36 encoder linears + 3 stem convs, augmentation1 (identity) pass")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 4                                                   # bounded sample of the B=64 workload (about 1 s of host time per step)
    warm = min(args.warmup, 1)
    sps, cores, sec = cpu_train_steps(B, args.labels, args.steps, warm, args.eeg_ch, lora_dropout=args.lora_dropout)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.eeg_ch, args.batch, args.labels, args.lora_dropout), "parallelism": "host cores",
                   "sample": f"each step is a bounded sample of that workload: B={B} instead of {args.batch} per step"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} training steps of B={B} (oracle port of the reference path; PEFT/accelerate absent)"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from neuspeech1_b200 import _abi, ops
    from neuspeech1_b200.engine import ModelDims
    from neuspeech1_b200.load_model import WhisperEEGForConditionalGeneration
    from neuspeech1_b200.weights import random_lora, random_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dims = ModelDims(eeg_ch=args.eeg_ch)
    B, L = args.batch, args.labels
    model = WhisperEEGForConditionalGeneration(dims, random_params(dims, seed=0), random_lora(dims, seed=1, b_std=0.01),
                                               dtype=torch.bfloat16, device=dev, lora_dropout=args.lora_dropout)
    model.train()
    eng = model.engine
    x_host, labels_host = synthetic_host_batch(dims, B, L, seed=100 + rank)
    x_host = x_host.pin_memory(); labels_host = labels_host.pin_memory()
    x_dev = x_host.to(dev); labels_dev = labels_host.to(dev)
    lr = 1e-3

    def allreduce(flat):
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / world)

    def step_resident():
        return eng.train_step(x_dev, labels_dev, lr=lr, all_reduce=allreduce if world > 1 else None)

    # e2e: host (pinned) batch -> H2D on a copy stream (double buffered, prefetching step i+1 during step i) -> training step
    # through the module API -> D2H of the loss (read back one step later so the host never stalls the launch queue).
    copy_stream = torch.cuda.Stream(device=dev)
    stages = [(torch.empty_like(x_dev), torch.empty_like(labels_dev)) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    ev_loss = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "losses": []}

    def h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[slot])
            stages[slot][0].copy_(x_host, non_blocking=True)
            stages[slot][1].copy_(labels_host, non_blocking=True)
            ev_ready[slot].record(copy_stream)

    def step_e2e():
        i = e2e_state["i"]
        slot = i & 1
        cur = torch.cuda.current_stream()
        if i == 0:
            ev_free[0].record(cur); ev_free[1].record(cur)
            h2d(0)
        h2d(slot ^ 1)                                          # prefetch the next step's batch
        cur.wait_event(ev_ready[slot])
        out = model.training_step(stages[slot][0], stages[slot][1], lr=lr, all_reduce=allreduce if world > 1 else None)
        ev_free[slot].record(cur)
        loss_host[slot].copy_(out.loss, non_blocking=True)     # D2H of the step's result
        ev_loss[slot].record(cur)
        if i > 0:
            ev_loss[slot ^ 1].synchronize()
            e2e_state["losses"].append(float(loss_host[slot ^ 1]))
        e2e_state["i"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    _abi.reset_counters()
    replayed0 = eng.graph_launches
    with ClockSampler(local) as clk:
        ms = timed(step_resident, args.steps)
    # launches issued one by one (counted in the library) + launches executed by CUDA-graph replays of pack/forward/backward
    # (counted once at capture, added per replay)
    launches = sum(_abi.counters().values()) + (eng.graph_launches - replayed0)
    clocks = clk.summary()
    ms_step = ms / args.steps
    value = world * B * 1e3 / ms_step

    for _ in range(max(args.warmup, 4)):                       # both staging slots seen twice: their graphs are captured here
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    assert all(l == l for l in e2e_state["losses"]), "NaN loss in the e2e run"
    e2e_value = world * B * 1e3 / ms_e2e

    line = None
    # ---- per-kernel profile of one step (CUDA events around every ns_* launch on the launching stream).  Every rank runs the
    # step (it contains the gradient all-reduce, a collective); only rank 0 keeps the record.
    ops.profile_begin()
    step_resident()
    prof = ops.profile_end()
    if rank == 0:
        peaks = measured_peaks()
        fam = {}
        for r in prof:
            f = fam.setdefault(r["name"], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
            f["ms"] += r["ms"]; f["flops"] += r["flops"]; f["bytes"] += r["bytes"]; f["n"] += 1
        tot = sum(f["ms"] for f in fam.values())
        top_name, top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/), null when not captured
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01c_traffic.json")))
            traffic = tj.get(top_name, {}).get("dram_bytes_per_launch_avg")
        except Exception:
            traffic = None
        if top["flops"] > 0:
            ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": top_name, "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tf_sustained"], "traffic": traffic, "launches": top["n"], "share_of_step": top["ms"] / tot,
                    "peak_source": peaks["src"] + " (sustained bf16: kernel timed inside a long step)"}
        else:
            ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": top_name, "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                    "traffic": traffic, "launches": top["n"], "share_of_step": top["ms"] / tot, "peak_source": peaks["src"]}
        step_tf = value / world * FLOP_PER_SAMPLE.get(args.eeg_ch, 267.92e9) / 1e12
        shares = {k: round(v["ms"] / tot, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:8]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sps, cores, sec = cpu_train_steps(4, L, 2, 1, args.eeg_ch, lora_dropout=args.lora_dropout)
            cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": "2 training steps of B=4 after 1 warm-up (oracle port of the reference path, fp32, all host threads)"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(args.eeg_ch, B, L, args.lora_dropout),
                       "parallelism": f"dp{world}", "l2": "inputs and activations per step (>10 GB) exceed the 126 MB L2",
                       "launch": "pack + forward + backward replayed as one CUDA graph per input buffer, all-reduce and optimizer launched eagerly"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": x_host.numel() * 4 + labels_host.numel() * 8,
                    "d2h_bytes_per_step": 4},
            "roofline": roof,
            "step_roofline": {"achieved_tflops": step_tf, "frac_of_sustained_peak": step_tf / peaks["tf_sustained"],
                              "flop_per_sample": FLOP_PER_SAMPLE.get(args.eeg_ch)},
            "kernel_shares": shares,
            "cpu_baseline": cpu,
        }
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"families": fam, "launches": prof}, f)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--labels", type=int, default=32)
    ap.add_argument("--eeg-ch", dest="eeg_ch", type=int, default=208)
    ap.add_argument("--lora-dropout", dest="lora_dropout", type=float, default=0.05)     # finetune.py:210
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
