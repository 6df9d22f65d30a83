#!/usr/bin/env python
"""Headline benchmark: THUMOS14 OpenTAL training clips/s on N B200s (one process per GPU, NCCL over NVLink).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--config thumos|anet] [--frames T] [--mode train|infer]
                    [--precision bf16x3|bf16] [--impl reference]

A step = one full training pass over one batch of synthetic clips: BDNet forward (native I3D backbone + head), the
MultiSegmentLoss (EDL + actionness + regression terms) and the boundary BCE of train.py, backward, gradient all-reduce
(N > 1) and the fused Adam update.  Workload = BASELINE.json configs[1]/[2]: configs/thumos14_opental_final.yaml
--open_set, clips 3x256x96x96, batch 8 per GPU, keyed synthetic weights, synthetic uint8 clips/targets/score maps
 — the default.  `--config anet` = configs[3] (768-frame clips, 150 classes), `--frames T --batch B` = one point of the
clip-length x batch sweep (configs[4]), `--mode infer` = BDNet forward only under the reference's timing protocol.  Prints ONE
JSON line (rank 0) — contract in the round prompt; `--impl reference` times the reference's own modules (oracle/_ref/reference_src,
else the oracle restatement) on the host cores instead, same config, same batch.
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

UNIT = "clips/s"
# conv FLOPs per clip at 256 frames, fp32 semantics: forward 168.43 G, fwd + dgrad + wgrad 466.45 G; linear
# in the clip length (backbone 0.638 GF/frame; the head is sized for T/4 positions)
FLOP_FWD_256 = 168.43e9
FLOP_TRAIN_256 = 466.45e9


def workload(args) -> dict:
    """The synthetic workload a `--config / --frames / --mode` combination names (BASELINE.json configs[1]-[4])."""
    anet = args.config == "anet"
    T = 768 if anet else args.frames
    mode = args.mode
    name = ("ActivityNet OpenTAL (configs/anet_opental.yaml --open_set)" if anet
            else "THUMOS14 OpenTAL (configs/thumos14_opental_final.yaml --open_set)")
    what = ("training step: BDNet fwd + MultiSegmentLoss(edl, IBM, actionness) + boundary BCE + bwd + Adam" if mode == "train"
            else "inference: BDNet forward (eval, no_grad), AFSD/thumos14/BDNet.py:564-583 protocol")
    metric = (f"{'training' if mode == 'train' else 'inference'} clips/sec ({'ActivityNet' if anet else 'THUMOS14'} OpenTAL, "
              f"3x{T}x96x96 clips)")
    return dict(anet=anet, frames=T, classes=150 if anet else 15, mode=mode, metric=metric,
                workload=f"{name} {what}; clips 3x{T}x96x96",
                flop_clip=(FLOP_TRAIN_256 if mode == "train" else FLOP_FWD_256) * T / 256.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU per step")
    ap.add_argument("--config", default="thumos", choices=["thumos", "anet"], help="thumos = BASELINE configs[1]/[2] (headline); "
                    "anet = configs[3]: 768-frame clips, 150 classes, configs/anet_opental.yaml")
    ap.add_argument("--frames", type=int, default=256, help="clip length of the thumos workload (configs[4]: 128..1024)")
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="infer = BDNet forward only (the reference's "
                    "published timing protocol, AFSD/thumos14/BDNet.py:564-583)")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the extra legs of the default run (ActivityNet config, "
                    "clip length x batch sweep, inference protocol: BASELINE.json configs[3], configs[4])")
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
# CPU baseline / reference arm: the REFERENCE's own modules (oracle/_ref/reference_src, placed by oracle/build_ref.py) on the
# host cores; the oracle restatement (oracle/opental_oracle.py) when they are not there.  The reference's CUDA-only
# BoundaryMaxPooling extension has no CPU form: oracle/ref_loader.py stands in for those 3 calls per forward (< 1 % of the time).
# ------------------------------------------------------------------------------------------------------------------
class CpuArm:
    """One process-wide instance: builds the CPU model once, then times steps of it."""

    def __init__(self, wl: dict):
        import torch
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import opental_oracle as O          # bench.py's cpu_baseline / reference arm is allowed to execute oracle/
        self.O, self.torch, self.wl = O, torch, wl
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        T = wl["frames"]
        self.cfg = O.anet_config() if wl["anet"] else O.OracleConfig(frame_num=T, feat_t=T // 4, clip_length=T)
        self.kind, self.ns = "port", None
        try:
            if not wl["anet"] and T != 256:
                # the reference's THUMOS14 model is built for its config's 256-frame clips (BDNet.py:18-22 read module globals):
                # the other clip lengths of the sweep run through the restatement, which takes the length as a parameter
                raise RuntimeError("reference model is fixed at 256 frames")
            import ref_loader
            self.ref_loader = ref_loader
            if wl["anet"]:
                self.ns = ref_loader.load_reference(config="configs/anet_opental.yaml", extra_args=("--open_set", "--split=0"), flavour="anet")
            else:
                self.ns = ref_loader.load_reference()
            self.kind = "reference"
        except Exception as e:  # noqa: BLE001 - no reference sources on this box: the restatement is the baseline
            self.why_port = repr(e)[:200]
        sd = O.synthetic_state_dict(self.cfg)
        if self.ns is not None:
            kw = dict(frame_num=768) if wl["anet"] else {}
            self.net = self.ns.BDNet(in_channels=3, training=False, use_edl=True, **kw)
            self.net.load_state_dict(sd)
            self.net.train()
            cfgt = self.ns.config["training"]
            kwl = {} if wl["anet"] else dict(act_config=cfgt["act_config"])
            self.crit = self.ns.MultiSegmentLoss(self.cfg.num_classes, 0.5, 1.0, cls_loss_type="edl", edl_config=cfgt["edl_config"],
                                                 os_head=True, **kwl)
            self.crit.cls_loss.epoch = 11
            self.opt = torch.optim.Adam(self.net.parameters(), lr=1e-5, weight_decay=1e-3)      # thumos14/train.py:321-323
        else:
            self.sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".bn." not in k else v)
                       for k, v in sd.items()}
            self.state = O.LossState(epoch=11)
            self.opt = torch.optim.Adam([v for v in self.sd.values() if v.requires_grad], lr=1e-5, weight_decay=1e-3)

    def batch(self, B: int):
        O, torch, wl = self.O, self.torch, self.wl
        x = torch.stack([O.synthetic_clip(i, frames=wl["frames"]) for i in range(B)])
        tg = [O.synthetic_targets(i, num_classes=wl["classes"]) for i in range(B)]
        sc = torch.stack([O.synthetic_scores(t, frames=wl["frames"]) for t in tg])
        return x, tg, sc

    def train_step(self, x, tg, sc) -> None:
        """forward + loss + backward + Adam, thumos14/train.py:226-252 (anet/train.py:168-230 for the ActivityNet flavour)."""
        O, wl = self.O, self.wl
        self.opt.zero_grad()
        if self.ns is not None and not wl["anet"]:
            cost, *_ = self.ref_loader.reference_training_cost(self.ns, self.net, self.crit, x, tg, sc)
        elif self.ns is not None:
            out = self.net(x)
            l = self.crit([out[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "priors", "act", "prop_act")], [t.clone() for t in tg])
            cost = l[0] + 10 * l[1] + l[2] + 10 * l[3] + l[4] + l[5] + l[6]
        elif wl["anet"]:
            out = O.bdnet_forward(x, self.sd, self.cfg, compat=True)
            l = O.multisegment_loss_anet(out, tg, self.state, self.cfg)
            cost = l[0] + 10 * l[1] + l[2] + 10 * l[3] + l[4] + l[5] + l[6]
        else:
            out = O.bdnet_forward(x, self.sd, self.cfg, compat=True)
            cost, _ = O.training_cost(out, tg, sc, self.state, self.cfg)
        cost.backward()
        self.opt.step()

    def forward(self, x) -> None:
        with self.torch.no_grad():
            if self.ns is not None:
                self.net(x)
            else:
                self.O.bdnet_forward(x, self.sd, self.cfg, compat=True)

    def time_steps(self, steps: int, warmup: int, B: int, mode: str):
        """(clips/s, ms/step): `steps` timed steps of batch B after `warmup` untimed ones."""
        x, tg, sc = self.batch(B)
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            if mode == "train":
                self.train_step(x, tg, sc)
            else:
                self.forward(x)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        ms = 1000.0 * sum(times) / len(times)
        return B * 1000.0 / ms, ms

    def describe(self) -> str:
        if self.kind == "reference":
            return ("the reference's own AFSD modules (BDNet, MultiSegmentLoss, torch.optim.Adam) on the host cores, torch CPU fp32; "
                    "BoundaryMaxPooling (CUDA-only in the reference) through the CPU stand-in of oracle/ref_loader.py")
        return "CPU restatement of the reference (oracle/opental_oracle.py, torch CPU fp32): reference sources not on this box"


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
    """`--impl reference`: the reference's CPU path on the SAME config (batch per step, optimizer included), K timed steps after W
    untimed ones as asked.  A CPU step of 8 clips takes several seconds, so the batch is what bounds the run: when K + W steps of
    the full batch would exceed ~4 minutes the per-step batch is reduced (and reported) — clips/s barely depends on it on a CPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args)
    arm = CpuArm(wl)
    B = args.batch
    # probe: one clip, to size the run (also warms the allocator / thread pool)
    v1, ms1 = arm.time_steps(1, 1, 1, wl["mode"])
    budget_s = 240.0
    while B > 1 and (args.steps + args.warmup) * B * ms1 / 1000.0 > budget_s:
        B //= 2
    try:
        val, ms = arm.time_steps(args.steps, args.warmup, B, wl["mode"])
    except (RuntimeError, MemoryError):          # host memory: halve the batch once more
        B = max(1, B // 2)
        val, ms = arm.time_steps(args.steps, args.warmup, B, wl["mode"])
    line = {
        "impl": "reference", "metric": wl["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
        "data": "synthetic",
        "config": {"workload": wl["workload"], "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
                   "parallelism": f"dp{args.gpus}", "reference_batch_per_step": B,
                   "reference_sample": f"{arm.describe()}; one step = one {'training step (optimizer included)' if wl['mode'] == 'train' else 'forward'} "
                                       f"of {B} clip(s), rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, **host_info(),
                         "sample": f"{args.steps} steps of {B} clip(s) after {args.warmup} warm-up"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(wl: dict) -> dict:
    """The `cpu_baseline` object of the native arm: a bounded sample (~10-30 s) of the same workload on the host cores."""
    arm = CpuArm(wl)
    b = 2 if wl["frames"] <= 256 else 1
    n = 4 if wl["mode"] == "train" else 6
    v, cms = arm.time_steps(n, 1, b, wl["mode"])
    return {"value": v, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, **host_info(), "ms_per_step": cms,
            "sample": f"{n} {'training steps (forward + loss + backward + Adam)' if wl['mode'] == 'train' else 'forwards'} of {b} clip(s) "
                      f"after 1 warm-up: {arm.describe()}"}


# activation bytes of one I3D forward per clip at 256 frames: every conv / pool output written once as two bf16 planes
# (4 B per element) and read by its consumers.  Elements per clip: Conv3d_1a 18.9 M, pool2a 4.7 M, 2b 4.7 M, 2c 14.2 M,
# pool3a 3.5 M, Mixed_3b 4.7 + 2.1 (bottlenecks) + 3.5 (pooled) M, Mixed_3c 8.8 + 2.9 + 4.7 M, pool4a 1.1 M, Mixed_4b..4f
# 5 x (1.2 + 0.3 + 1.2) M, pool5a 0.24 M, Mixed_5b/5c 2 x (0.3 + 0.06 + 0.24) M = ~89 M elements.
ACT_ELEMS_256 = 89e6


def hbm_estimate(wl: dict, B: int, ms: float, peaks: dict) -> dict:
    """HBM roofline of the whole step (BASELINE configs[4] asks for both rooflines): an ALGORITHMIC lower bound of the bytes a
    step must move — every activation written once and read once in the forward (2 x 4 B per element), and in the backward read
    again by the weight-gradient and ReLU kernels, with a gradient tensor of the same size written and read (4 x 4 B) — against
    the measured copy bandwidth.  Weights (12 M parameters) and the 1-D head are negligible next to that."""
    per_elem = 8.0 if wl["mode"] == "infer" else 24.0
    nbytes = ACT_ELEMS_256 * wl["frames"] / 256.0 * per_elem * B
    peak = peaks.get("hbm_gbs") or 6455.9
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "algorithmic_bytes_per_step": nbytes,
            "note": "whole-step lower bound: activation planes written / read once per producer / consumer; the step is tensor-bound "
                    "in its 3x3x3 convolutions and HBM-bound in its 1x1 convolutions, pools and elementwise kernels"}


def run_native_infer(args, wl, net, dev, world, rank, local):
    """`--mode infer`: BDNet forward only.  (1) the reference's published protocol (AFSD/thumos14/BDNet.py:564-583): random fp32
    input [1,3,T,96,96], 2 warm-ups, mean wall time of N synchronised calls -> ms and infer_fps; (2) throughput at batch B from
    uint8 frames resident in HBM, forward replayed as a CUDA graph, CUDA events; (3) end to end from pinned host frames with the
    head outputs read back."""
    import torch
    import torch.distributed as dist

    from opental_b200 import _lib, ops
    B, T = args.batch, wl["frames"]
    net.eval()
    keys = ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "unct", "prop_unct")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # (1) reference protocol
        x1 = torch.randn(1, 3, T, 96, 96, device=dev)
        for _ in range(2):
            net(x1)
        torch.cuda.synchronize()
        n1 = max(5, min(args.steps, 50))
        t_run = 0.0
        for _ in range(n1):
            torch.cuda.synchronize(); t0 = time.time()
            net(x1)
            torch.cuda.synchronize(); t_run += time.time() - t0
        ms_b1 = t_run / n1 * 1e3
        # (2) batch B, uint8 frames, graph replay
        from opental_b200 import engine
        host = [torch.stack([engine.synthetic_clip_u8(rank * 1000 + j * B + i, rank, frames=T) for i in range(B)]).pin_memory() for j in range(2)]
        static = host[0].to(dev)
        devb = [h.to(dev) for h in host]
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            for _ in range(2):
                net(static)
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
            out = net(static)
        launches_per_graph = _lib.launch_count() - n0

        def step(j, src=None):
            static.copy_(devb[j % 2] if src is None else src, non_blocking=True)
            graph.replay()

        for j in range(args.warmup):
            step(j)
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        for j in range(args.steps):
            step(j)
        e1.record()
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1) / args.steps
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
        value = world * B * 1000.0 / ms
        # (3) end to end
        e2e = None
        if not args.no_e2e:
            outs_host = {k: torch.empty_like(out[k], device="cpu").pin_memory() for k in keys if out.get(k) is not None}
            barrier()
            t0 = time.perf_counter()
            for j in range(args.steps):
                step(j, host[j % 2])
                for k, h in outs_host.items():
                    h.copy_(out[k], non_blocking=True)
                torch.cuda.synchronize()
            ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
            if world > 1:
                t = torch.tensor([ms_e2e], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_e2e = float(t)
            e2e = {"value": world * B * 1000.0 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": host[0].numel(),
                   "d2h_bytes_per_step": sum(h.numel() * 4 for h in outs_host.values()), "ms_per_step": ms_e2e,
                   "api": f"opental_b200.bdnet.BDNet.forward (eval, no_grad) on pinned host uint8 frames [B,{T},112,112,3]; the head "
                          "outputs (loc, conf, prop_*, center, act, unct) are copied back every step"}
        # per-kernel roofline: eager forward with the tensor-core launches bracketed by events
        with ops.PROFILE.enabled() as prof:
            for j in range(3):
                net(devb[j % 2])
            barrier()
        ksum = prof.summary()
    if rank != 0:
        shutdown()
        return
    peaks = load_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    hw = 3.0 if args.precision == "bf16x3" else 1.0
    roof = {}
    for k, d in ksum.items():
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        roof[k] = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                   "launches_per_step": d["launches"] / 3, "ms_per_step": d["ms"] / 3, "share_of_step": d["ms"] / 3 / ms,
                   "executed_tflops": ach * hw, "executed_frac": ach * hw / peak_tf}
    big = [k for k in roof if "head" not in k]
    dominant = max(big, key=lambda k: roof[k]["ms_per_step"]) if big else None
    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (3 bf16 tensor-core passes, fp32 accumulate; fp32-equivalent ~1e-5)" if args.precision == "bf16x3" else "bf16",
        "data": "synthetic",
        "config": {"workload": wl["workload"], "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "precision": args.precision, "input": f"uint8 frames [B,{T},112,112,3]", "cuda_graph": True,
                   "l2": "activations of a batch-8 forward (~3 GB) exceed the 126 MB L2; two alternating input batches"},
        "reference_protocol": {"input": f"randn [1,3,{T},96,96] fp32", "warmup": 2, "runs": n1, "ms": ms_b1, "infer_fps": 1000.0 / ms_b1,
                               "how": "eager call bracketed by torch.cuda.synchronize(), wall clock, mean (AFSD/thumos14/BDNet.py:564-583)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_graph * args.steps,
        "roofline": roof.get(dominant), "roofline_kernel": dominant, "roofline_all": roof,
        "model_tflops_per_gpu": value / world * wl["flop_clip"] / 1e12,
        "hbm": hbm_estimate(wl, B, ms, peaks),
        "cpu_baseline": None if args.no_cpu_baseline else cpu_baseline(wl),
    }
    print(json.dumps(line), flush=True)
    shutdown()


def shutdown(objs=()) -> None:
    """Leave the process group without hanging (engine.shutdown_distributed); the JSON line is out by then."""
    from opental_b200 import engine
    sys.stdout.flush()
    engine.shutdown_distributed(objs)


def load_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:  # noqa: BLE001
        return {}


# ------------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configs, measured in the same run: configs[3] (ActivityNet) and configs[4] (clip length x batch)
# ------------------------------------------------------------------------------------------------------------------
def measure_point(tr, *, anet: bool, T: int, B: int, dev, world: int, rank: int, peaks: dict, steps: int = 5, warmup: int = 2) -> dict:
    """One point of the sweep on an existing Trainer: a CUDA-graph captured training step at batch B of T-frame uint8 clips
    resident in HBM, `warmup` + `steps` replays, CUDA events, MAX over ranks (barrier + synchronize on both sides); both
    rooflines (algorithmic conv FLOPs vs the measured bf16 peak; algorithmic activation bytes vs the measured copy bandwidth)."""
    import torch
    import torch.distributed as dist

    from opental_b200 import engine
    from opental_b200.multisegment_loss import pad_targets
    torch.cuda.reset_peak_memory_stats()
    g = torch.Generator(device=dev).manual_seed(1000 * rank + B)
    clips = torch.randint(0, 256, (B, T, 112, 112, 3), generator=g, dtype=torch.uint8, device=dev)     # i.i.d. pixels
    tg = [engine.synthetic_targets(i, rank, num_classes=150 if anet else 15) for i in range(B)]
    sc = torch.stack([engine.synthetic_scores(t, frames=T) for t in tg]).to(dev)
    tp, tv = (t.to(dev) for t in pad_targets(tg, device="cpu"))
    tr._graph = None
    mode = "graph"
    try:
        tr.capture(clips, (tp, tv), sc)
    except Exception:  # noqa: BLE001   (B x priors > 4096: the loss takes its torch formulation, which synchronises)
        tr._graph = None
        mode = "eager"
        torch.cuda.synchronize()
    try:
        for _ in range(warmup):
            tr.step(clips, (tp, tv), sc)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.step(clips, (tp, tv), sc)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    finally:
        tr._graph = tr._graph_out = tr._static = None
        tr._graph_cache.clear()
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    tf = B * FLOP_TRAIN_256 * T / 256.0 / (ms * 1e-3) / 1e12
    hbm = hbm_estimate(dict(mode="train", frames=T), B, ms, peaks)
    return {"frames": T, "batch_per_gpu": B, "ms_per_step": round(ms, 3), "clips_per_s": round(world * B * 1000.0 / ms, 2),
            "tensor_tflops_per_gpu": round(tf, 1), "tensor_frac": round(tf / peak_tf, 4), "hbm_gbs": round(hbm["achieved"], 0),
            "hbm_frac": round(hbm["frac"], 4), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1), "mode": mode}


def other_configs(args, dev, world: int, rank: int, peaks: dict) -> dict:
    """BASELINE.json configs[3] and configs[4] and the inference protocol on the SAME GPUs in the SAME run, a few steps each
    (bounded: ~30 s), so that the driver's 1/2/4/8-GPU runs of the default command carry them.  Every rank takes part (the
    training points contain the gradient all-reduce).  A failing point is reported as text and does not stop the line."""
    import torch

    from opental_b200 import engine
    out = {"how": "each point: CUDA-graph captured training step, 2 warm-up + 5 timed replays, CUDA events, max over ranks; uint8 "
                  "clips resident in HBM; clips_per_s is the whole job's; tensor_frac = algorithmic conv FLOPs / measured sustained "
                  "bf16 peak, hbm_frac = algorithmic activation bytes / measured copy bandwidth (both per GPU)"}

    def release():
        import gc
        gc.collect()
        torch.cuda.empty_cache()

    # configs[3]: ActivityNet OpenTAL, 768-frame clips, 150 classes, per-sample loss, backbone at 0.1 x the head's rate
    try:
        torch.manual_seed(0)
        net, crit = engine.build_opental_anet(device=dev, precision=args.precision, epoch=11)
        tr = engine.Trainer(net, crit, lr=1e-5, weight_decay=1e-3, backbone_lr_scale=0.1)
        tr.broadcast_parameters(0)
        out["anet"] = {"workload": "ActivityNet OpenTAL (configs/anet_opental.yaml --open_set) training step; clips 3x768x96x96",
                       **measure_point(tr, anet=True, T=768, B=args.batch, dev=dev, world=world, rank=rank, peaks=peaks)}
        del tr, net, crit
    except Exception as ex:  # noqa: BLE001
        out["anet"] = {"error": repr(ex)[:200]}
    release()
    # configs[4]: clip length 128..1024 x batch 1..16
    pts = []
    for T in (128, 256, 512, 1024):
        try:
            torch.manual_seed(0)
            net, crit = engine.build_opental(device=dev, precision=args.precision, frame_num=T, epoch=11)
            tr = engine.Trainer(net, crit, lr=1e-5, weight_decay=1e-3)
            tr.broadcast_parameters(0)
            for B in (1, 2, 4, 8, 16):
                try:
                    pts.append(measure_point(tr, anet=False, T=T, B=B, dev=dev, world=world, rank=rank, peaks=peaks))
                except Exception as ex:  # noqa: BLE001
                    pts.append({"frames": T, "batch_per_gpu": B, "error": repr(ex)[:200]})
                release()
            del tr, net, crit
        except Exception as ex:  # noqa: BLE001
            pts.append({"frames": T, "error": repr(ex)[:200]})
        release()
    out["cliplen_batch_sweep"] = pts
    # inference under the reference's only published protocol (AFSD/thumos14/BDNet.py:564-583) + batch-8 forward throughput
    try:
        torch.manual_seed(0)
        net, _ = engine.build_opental(device=dev, precision=args.precision, frame_num=256, epoch=11)
        net.eval()
        with torch.no_grad():
            x1 = torch.randn(1, 3, 256, 96, 96, device=dev)
            for _ in range(2):
                net(x1)
            torch.cuda.synchronize()
            t_run = 0.0
            for _ in range(10):
                torch.cuda.synchronize(); t0 = time.time()
                net(x1)
                torch.cuda.synchronize(); t_run += time.time() - t0
            ms1 = t_run / 10 * 1e3
            xb = torch.randint(0, 256, (args.batch, 256, 112, 112, 3), dtype=torch.uint8, device=dev)
            for _ in range(2):
                net(xb)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                net(xb)
            e1.record()
            torch.cuda.synchronize()
            msb = e0.elapsed_time(e1) / 5
        out["inference"] = {"reference_protocol_ms": round(ms1, 3), "infer_fps": round(1000.0 / ms1, 1),
                            "protocol": "randn [1,3,256,96,96], 2 warm-ups, mean wall time of 10 synchronised eager forwards",
                            "batch": args.batch, "forward_ms": round(msb, 3), "forward_clips_per_s_per_gpu": round(args.batch * 1000.0 / msb, 1),
                            "forward_tensor_frac": round(args.batch * FLOP_FWD_256 / (msb * 1e-3) / 1e12 / (peaks.get("bf16_tflops_sustained") or 1400.0), 4),
                            "note": "eager forwards (host-enqueued); `--mode infer` gives the graph-replayed number with its e2e leg"}
        del net
    except Exception as ex:  # noqa: BLE001
        out["inference"] = {"error": repr(ex)[:200]}
    release()
    return out


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
    wl = workload(args)
    T, anet = wl["frames"], wl["anet"]

    torch.manual_seed(0)
    if anet:
        net, crit = engine.build_opental_anet(device=dev, precision=args.precision, epoch=11)
    else:
        net, crit = engine.build_opental(device=dev, precision=args.precision, frame_num=T, epoch=11)
    # keyed synthetic weights would need the oracle; the bench uses the module's own deterministic init (same shapes)
    if wl["mode"] == "infer":
        return run_native_infer(args, wl, net, dev, world, rank, local)
    # anet/train.py:303-310 trains the backbone at 0.1 x the head's rate
    tr = engine.Trainer(net, crit, lr=1e-5, weight_decay=1e-3, backbone_lr_scale=0.1 if anet else 1.0)
    tr.broadcast_parameters(0)

    # two distinct synthetic batches per rank, host (pinned) and device copies.  Clips travel in the dataset's storage
    # format — uint8 frames [B,T,112,112,3] (AFSD/common/video2npy.py:61-74) — and are cropped to 96x96 and normalised
    # to [-1,1] by the ingest kernel on the device (thumos_dataset.py:261-263), not by a CPU loader.
    def make_batch(j):
        idx = [rank * 1000 + j * B + i for i in range(B)]
        clips = torch.stack([engine.synthetic_clip_u8(i, rank, frames=T) for i in idx])
        tg = [engine.synthetic_targets(i, rank, num_classes=wl["classes"]) for i in idx]
        sc = torch.stack([engine.synthetic_scores(t, frames=T) for t in tg])
        from opental_b200.multisegment_loss import pad_targets
        tp, tv = pad_targets(tg, device="cpu")
        out = [clips.pin_memory(), tp.pin_memory(), tv.pin_memory(), sc.pin_memory()]
        if args.ssl:
            # cut-paste plan per clip (thumos_dataset.py:187-237); th = 8 frames; re-draw until the attempt succeeds
            import random
            from opental_b200 import augment
            maps, ssl_tg = [], []
            for i, t in zip(idx, tg):
                annos = [[float(a) * T, float(b) * T, int(c)] for a, b, c in t.tolist()]
                rng, flag = random.Random(i), False
                while not flag:
                    fmap, new_annos, flag = augment.cut_paste(annos, 8, T, 1, rng=rng)
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
    graph_note = None
    if not args.no_graph:
        n0 = _lib.launch_count()
        ssl_kw = dict(ssl_targets=list(devb[0][5].unbind(0)), ssl_frame_map=devb[0][4]) if args.ssl else {}
        try:
            tr.capture(devb[0][0], (devb[0][1], devb[0][2]), devb[0][3], **ssl_kw)
            launches_per_graph = (_lib.launch_count() - n0) // 3      # capture() runs the step 3x (2 warm-ups + the capture)
        except Exception as ex:  # noqa: BLE001   (B x priors > 4096: the loss takes its torch formulation, which synchronises)
            tr._graph = None
            args.no_graph = True
            graph_note = repr(ex)[:160]
            torch.cuda.synchronize()
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
        # the loss of every step is read back on the host, one step behind the GPU (what a training loop that logs its loss does):
        # step j's cost goes to a pinned slot with an async copy, the host reads it while step j + 1 runs
        cost_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        cost_ready = [torch.cuda.Event(), torch.cuda.Event()]
        seen = []
        barrier()
        t0 = time.perf_counter()
        prefetch(0)
        for j in range(args.steps):
            if j + 1 < args.steps:
                prefetch(j + 1)
            torch.cuda.current_stream().wait_event(ready[j % 2])
            cost, losses, ls, le = run_step(bufs[j % 2])
            consumed[j % 2].record()
            cost_host[j % 2:j % 2 + 1].copy_(cost.reshape(1), non_blocking=True)      # device -> host read of the step's result (4 bytes)
            cost_ready[j % 2].record()
            if j >= 1:
                cost_ready[(j - 1) % 2].synchronize()
                seen.append(float(cost_host[(j - 1) % 2]))
        cost_ready[(args.steps - 1) % 2].synchronize()
        seen.append(float(cost_host[(args.steps - 1) % 2]))
        barrier()
        assert len(seen) == args.steps and all(v == v for v in seen), "every step's loss must have been read"
        ms_e2e = (time.perf_counter() - t0) * 1000.0 / args.steps
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t)
        e2e = {"value": world * B * 1000.0 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e, "api": f"opental_b200.engine.Trainer.step on pinned host uint8 frames [B,{T},112,112,3] (H2D prefetched on a copy stream); every step's cost is copied to pinned host memory and read by the host one step behind the GPU"}

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

    peaks = load_peaks()
    others = None
    default_run = not anet and T == 256 and not args.ssl and args.precision == "bf16x3" and not args.no_graph
    if default_run and not args.no_other_configs:
        # free the headline model first: the sweep's largest point (1024 frames x 16) wants ~43 GB
        del graph, devb
        tr._graph = tr._graph_out = tr._static = None
        tr._graph_cache.clear()
        others = other_configs(args, dev, world, rank, peaks)

    if rank != 0:
        shutdown([tr])
        return

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
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as fh:
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
        cpu_base = cpu_baseline(wl)

    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (3 bf16 tensor-core passes, fp32 accumulate; fp32-equivalent ~1e-5)" if args.precision == "bf16x3" else "bf16",
        "data": "synthetic",
        "config": {"workload": wl["workload"],
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}", "precision": args.precision,
                   "input": f"uint8 frames [B,{T},112,112,3], centre crop 96 + normalisation in the ingest kernel",
                   "l2": "per-step activations and gradients (several GB) exceed the 126 MB L2; two alternating input batches",
                   "ssl_pass": bool(args.ssl), "cuda_graph": not args.no_graph, **({"cuda_graph_note": graph_note} if graph_note else {}),
                   "conv1a": "bf16x3 on the normalised clip (OTAL_U8_CONV1A=0)" if os.environ.get("OTAL_U8_CONV1A") == "0" else "raw uint8 pixels, one exact bf16 plane",
                   "staged_switches": sorted(k for k in ("OTAL_U8_CONV1A", "OTAL_FUSE_B12A", "OTAL_HEAD_SCHEDULE", "OTAL_CONV1A_WGRAD_HALO", "OTAL_NO_NCAT",
                                                         "OTAL_NO_WGRAD_OVERLAP") if os.environ.get(k))},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "dp_params_in_sync": in_sync,
        "gpu_launches_by_entry_point": dict(_lib.LAUNCHES),
        "roofline": roof.get(dominant),
        "roofline_kernel": dominant,
        "roofline_all": roof,
        "model_tflops_per_gpu": value / world * wl["flop_clip"] / 1e12,
        "hbm": hbm_estimate(wl, B, ms, peaks),
        "cpu_baseline": cpu_base,
        "other_configs": others,
    }
    print(json.dumps(line), flush=True)
    shutdown([tr])


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
