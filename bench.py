#!/usr/bin/env python
"""Headline benchmark: THUMOS14 OpenTAL training clips/s on N B200s (one process per GPU, NCCL over NVLink).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision bf16x3|bf16] [--impl reference]

A step = one full training pass over one batch of synthetic clips: BDNet forward (native I3D backbone + head), the
MultiSegmentLoss (EDL + actionness + regression terms) and the boundary BCE of train.py, backward, gradient all-reduce
(N > 1) and the fused Adam update.  Workload = BASELINE.json configs[1]/[2]: configs/thumos14_opental_final.yaml
--open_set, clips 3x256x96x96, batch 8 per GPU, keyed synthetic weights, synthetic uint8 clips/targets/score maps
(SURVEY §8d).  Prints ONE JSON line (rank 0) — contract in the round prompt; `--impl reference` times the CPU
restatement of the reference (oracle/) on the host cores instead.
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
sys.path.insert(0, ROOT)

METRIC = "training clips/sec (THUMOS14 OpenTAL, 3x256x96x96 clips)"
WORKLOAD = ("THUMOS14 OpenTAL (configs/thumos14_opental_final.yaml --open_set) training step: BDNet fwd + MultiSegmentLoss(edl, "
            "IBM, actionness) + boundary BCE + bwd + Adam; clips 3x256x96x96")
UNIT = "clips/s"
FLOP_TRAIN_PER_CLIP = 466.45e9      # fwd + dgrad + wgrad conv FLOPs, fp32 semantics


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU per step")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ssl", action="store_true", help="add the self-supervised second pass (cut-paste clip through the frame map + "
                    "triplet loss, train.py:237-242) to every step; the headline number is quoted without it (SURVEY §8d)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every launch from the host instead of replaying the captured step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.3:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle restatement of the reference training step on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_training_steps(steps: int, warmup: int, batch: int = 1):
    """Times `steps` full training steps (forward, loss, backward; batch `batch`) of the CPU restatement of the
    reference (oracle/opental_oracle.py: torch CPU fp32 ops, all host threads).  Returns (clips/s, ms/step, cores)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import opental_oracle as O          # bench.py's cpu_baseline / reference arm is allowed to execute oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".bn." not in k else v)
          for k, v in sd.items()}
    state = O.LossState(epoch=11)
    x = torch.stack([O.synthetic_clip(i) for i in range(batch)])
    targets = [O.synthetic_targets(i, num_classes=cfg.num_classes) for i in range(batch)]
    scores = torch.stack([O.synthetic_scores(t) for t in targets])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.bdnet_forward(x, sd, cfg, compat=True)
        cost, _ = O.training_cost(out, targets, scores, state, cfg)
        cost.backward()
        for v in sd.values():
            if v.is_floating_point() and v.grad is not None:
                v.grad = None
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    return batch * 1000.0 / ms, ms, cores


def cpu_forward_steps(steps: int = 3, warmup: int = 1, batch: int = 1):
    """Forward-only counterpart of cpu_training_steps (the CPU baseline reports (i) forward and (ii) forward + loss
    + backward): clips/s of `O.bdnet_forward` under no_grad, median of `steps` runs."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import opental_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg)
    x = torch.stack([O.synthetic_clip(i) for i in range(batch)])
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            O.bdnet_forward(x, sd, cfg, compat=True)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    return batch / times[len(times) // 2]


def host_info() -> dict:
    """CPU model and torch version of the box the CPU numbers were taken on (SURVEY §8d 'CPU baseline timing')."""
    import torch
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.lower().startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cpu": model, "torch": torch.__version__}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))          # one step = one clip (~0.5 s on 16 cores): the whole run stays under a minute
    warm = max(0, min(args.warmup, 2))
    val, ms, cores = cpu_training_steps(steps, warm, batch=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
                   "parallelism": f"dp{args.gpus}",
                   "reference_sample": "CPU restatement of the reference (oracle/, torch CPU fp32, all host cores); one step = forward + "
                                       "loss + backward of ONE clip (no optimizer), rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", **host_info(),
                         "forward_only_clips_per_s": cpu_forward_steps(3, 1, 1),
                         "sample": f"{steps} training steps of 1 clip after {warm} warm-up (forward + loss + backward, torch CPU fp32)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    from opental_b200 import _lib, engine, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the native path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    torch.manual_seed(0)
    net, crit = engine.build_opental(device=dev, precision=args.precision, epoch=11)
    # keyed synthetic weights would need the oracle; the bench uses the module's own deterministic init (same shapes)
    tr = engine.Trainer(net, crit, lr=1e-5, weight_decay=1e-3)
    tr.broadcast_parameters(0)

    # two distinct synthetic batches per rank, host (pinned) and device copies.  Clips travel in the dataset's storage
    # format — uint8 frames [B,T,112,112,3] (AFSD/common/video2npy.py:61-74) — and are cropped to 96x96 and normalised
    # to [-1,1] by the ingest kernel on the device (thumos_dataset.py:261-263), not by a CPU loader.
    def make_batch(j):
        idx = [rank * 1000 + j * B + i for i in range(B)]
        clips = torch.stack([engine.synthetic_clip_u8(i, rank) for i in idx])
        tg = [engine.synthetic_targets(i, rank) for i in idx]
        sc = torch.stack([engine.synthetic_scores(t) for t in tg])
        from opental_b200.multisegment_loss import pad_targets
        tp, tv = pad_targets(tg, device="cpu")
        out = [clips.pin_memory(), tp.pin_memory(), tv.pin_memory(), sc.pin_memory()]
        if args.ssl:
            # cut-paste plan per clip (thumos_dataset.py:187-237); th = 8 frames; re-draw until the attempt succeeds
            import random
            from opental_b200 import augment
            maps, ssl_tg = [], []
            for i, t in zip(idx, tg):
                annos = [[float(a) * 256, float(b) * 256, int(c)] for a, b, c in t.tolist()]
                rng, flag = random.Random(i), False
                while not flag:
                    fmap, new_annos, flag = augment.cut_paste(annos, 8, 256, 1, rng=rng)
                maps.append(torch.from_numpy(fmap))
                ssl_tg.append(torch.tensor(new_annos, dtype=torch.float32))
            out += [torch.stack(maps).pin_memory(), torch.stack(ssl_tg).pin_memory()]
        return tuple(out)

    host = [make_batch(j) for j in range(2)]
    devb = [tuple(t.to(dev) for t in hb) for hb in host]

    def run_step(batch):
        c, tp, tv, sc = batch[:4]
        if args.ssl:
            return tr.step(c, (tp, tv), sc, ssl_targets=list(batch[5].unbind(0)), ssl_frame_map=batch[4])
        return tr.step(c, (tp, tv), sc)

    def step_dev(j):
        return run_step(devb[j % 2])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_graph = 0
    if not args.no_graph:
        n0 = _lib.launch_count()
        ssl_kw = dict(ssl_targets=list(devb[0][5].unbind(0)), ssl_frame_map=devb[0][4]) if args.ssl else {}
        tr.capture(devb[0][0], (devb[0][1], devb[0][2]), devb[0][3], **ssl_kw)
        launches_per_graph = (_lib.launch_count() - n0) // 3          # capture() runs the step 3x (2 warm-ups + the capture)
    for j in range(args.warmup):
        step_dev(j)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for j in range(args.steps):
        step_dev(j)
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count() - launches0 + launches_per_graph * args.steps
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = world * B * 1000.0 / ms

    # ---- end to end: pinned host buffers -> H2D every step (prefetched on a copy stream) -> step -> loss read back
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        bufs = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(j):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[j % 2])
                for d, h in zip(bufs[j % 2], host[j % 2]):
                    d.copy_(h, non_blocking=True)
                ready[j % 2].record(copy_stream)

        for ev in consumed:
            ev.record()
        h2d = sum(t.numel() * t.element_size() for t in host[0])
        barrier()
        t0 = time.perf_counter()
        prefetch(0)
        for j in range(args.steps):
            if j + 1 < args.steps:
                prefetch(j + 1)
            torch.cuda.current_stream().wait_event(ready[j % 2])
            cost, losses, ls, le = run_step(bufs[j % 2])
            consumed[j % 2].record()
            _ = float(cost)                                   # device -> host read of the step's result (4 bytes)
        barrier()
        ms_e2e = (time.perf_counter() - t0) * 1000.0 / args.steps
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t)
        e2e = {"value": world * B * 1000.0 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e, "api": "opental_b200.engine.Trainer.step on pinned host uint8 frames [B,256,112,112,3] (H2D prefetched on a copy stream) + float(cost)"}

    # ---- per-kernel roofline: the same step enqueued eagerly with every tensor-core conv launch bracketed by CUDA events
    # on the launching stream (a graph replay cannot be bracketed per kernel); same process, same buffers, after the timed
    # region, on every rank (the step contains the gradient all-reduce).  FLOPs are algorithmic (fp32 semantics).
    graph, tr._graph = tr._graph, None
    n_prof = max(1, min(args.steps, 3))
    with ops.PROFILE.enabled() as prof:
        for j in range(n_prof):
            step_dev(j)
        barrier()
    ksum = prof.summary()
    tr._graph = graph

    # data-parallel sanity: after identical updates every rank must hold bit-identical parameters
    in_sync = None
    if world > 1:
        chk = torch.stack([w.double().sum() for w, _ in tr.groups] + [w.double().abs().sum() for w, _ in tr.groups])
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        in_sync = bool(torch.equal(hi, lo))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks.get("bf16_tflops_sustained") else "fallback 1.4 PFLOP/s sustained"
    roof = {}
    for k, d in ksum.items():
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        hw = 3.0 if args.precision == "bf16x3" else 1.0
        roof[k] = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                   "traffic": None, "launches_per_step": d["launches"] / n_prof,
                   "ms_per_step": d["ms"] / n_prof, "share_of_step": d["ms"] / n_prof / ms,
                   "executed_tflops": ach * hw, "executed_frac": ach * hw / peak_tf,
                   "note": f"achieved = algorithmic fp32-semantic conv FLOPs / event-timed kernel time; {args.precision} executes {hw:.0f}x "
                           f"those FLOPs on the bf16 pipe; peak = {peak_src}"}
    try:      # DRAM traffic of the dominant kernel's largest launch, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as fh:
            ncu_traffic = json.load(fh)
    except Exception:  # noqa: BLE001
        ncu_traffic = {}
    for k in roof:
        t = ncu_traffic.get(k)
        if t:
            roof[k]["traffic"] = t["dram_bytes"]
            roof[k]["traffic_note"] = (f"dram__bytes_read+write of one launch ({t['launch']}): {t['dram_bytes'] / 1e6:.0f} MB vs "
                                       f"{t['algorithmic_bytes'] / 1e6:.0f} MB algorithmic; tensor pipe active {t['tensor_pipe_active_pct']}% (ncu)")
    big = [k for k in roof if "head" not in k]
    dominant = max(big, key=lambda k: roof[k]["ms_per_step"]) if big else None

    cpu_base = None
    if not args.no_cpu_baseline:
        v, cms, cores = cpu_training_steps(8, 1, batch=2)
        cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", **host_info(),
                    "forward_only_clips_per_s": cpu_forward_steps(3, 1, 1),
                    "sample": "8 training steps of 2 clips after 1 warm-up, ~10 s of CPU work (forward + loss + backward, torch CPU "
                              "fp32 restatement of the reference in oracle/, all host cores)",
                    "ms_per_step": cms}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (3 bf16 tensor-core passes, fp32 accumulate; fp32-equivalent ~1e-5)" if args.precision == "bf16x3" else "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}", "precision": args.precision,
                   "input": "uint8 frames [B,256,112,112,3], centre crop 96 + normalisation in the ingest kernel",
                   "l2": "per-step activations and gradients (several GB) exceed the 126 MB L2; two alternating input batches",
                   "ssl_pass": bool(args.ssl), "cuda_graph": not args.no_graph,
                   "conv1a": "bf16x3 on the normalised clip (OTAL_U8_CONV1A=0)" if os.environ.get("OTAL_U8_CONV1A") == "0" else "raw uint8 pixels, one exact bf16 plane",
                   "staged_switches": sorted(k for k in ("OTAL_U8_CONV1A", "OTAL_FUSE_B12A", "OTAL_CONV_KSPLIT", "OTAL_NO_NCAT",
                                                         "OTAL_NO_WGRAD_OVERLAP") if os.environ.get(k))},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "dp_params_in_sync": in_sync,
        "gpu_launches_by_entry_point": dict(_lib.LAUNCHES),
        "roofline": roof.get(dominant),
        "roofline_kernel": dominant,
        "roofline_all": roof,
        "model_tflops_per_gpu": value / world * FLOP_TRAIN_PER_CLIP / 1e12,
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
