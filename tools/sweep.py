"""Clip-length x batch sweep (BASELINE.json configs[4]): training clips/s of one GPU for T in {128,256,512,1024} and
B in {1,2,4,8,16}, with the per-point tensor-pipe roofline fraction (algorithmic conv FLOPs / measured bf16 peak).

    python tools/sweep.py [--frames 128 256 512 1024] [--batches 1 2 4 8 16] > gpurun_out/sweep.txt
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py [--anet] ...     # data parallel, weak scaling

Under torchrun every rank runs the sweep on its own GPU with the gradient all-reduce of `engine.Trainer`; a point's time is
the MAX over ranks (barrier + synchronize on both sides), clips/s is the whole job's, rank 0 prints.

Conv FLOPs scale linearly in T: backbone 0.638 GFLOP/frame forward; the head is sized for T/4 positions.
Each point: CUDA-graph captured step, 2 warm-up + 5 timed replays, CUDA events."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from opental_b200 import engine  # noqa: E402
from opental_b200.multisegment_loss import pad_targets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="+", default=[128, 256, 512, 1024])
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 2, 4, 8, 16])
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--anet", action="store_true", help="ActivityNet flavour (configs/anet_opental.yaml): 768-frame clips, 150 classes")
    args = ap.parse_args()
    peak = 1398.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:  # noqa: BLE001
        pass
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def say(line):
        if rank == 0:
            print(line, flush=True)

    import bench as B_
    peaks = B_.load_peaks()
    say(f"# frames batch/GPU ms/step clips/s(job) train_TFLOP/s(alg, per GPU) frac_of_{peak:.0f}TF  HBM_GB/s(alg lower bound) "
        f"frac_of_{peaks.get('hbm_gbs', 6455.9):.0f}GB/s  peak_mem_GB   ({args.precision}, {world} GPU)")
    if args.anet:
        args.frames = [768]
    for T in args.frames:
        torch.manual_seed(0)
        if args.anet:
            net, crit = engine.build_opental_anet(device=dev, precision=args.precision, epoch=11)
        else:
            net, crit = engine.build_opental(device=dev, precision=args.precision, frame_num=T, epoch=11)
        tr = engine.Trainer(net, crit)
        tr.broadcast_parameters(0)
        # fwd + dgrad + wgrad conv FLOPs per clip, linear in T (466.45 GF at T = 256)
        flop_clip = 466.45e9 * T / 256.0
        for B in args.batches:
            try:
                torch.cuda.reset_peak_memory_stats()
                torch.manual_seed(1000 * rank + B)
                clips = torch.rand(B, 3, T, 96, 96, device=dev) * 2 - 1
                tg = [engine.synthetic_targets(i, rank, num_classes=150 if args.anet else 15) for i in range(B)]
                sc = torch.stack([engine.synthetic_scores(t, frames=T) for t in tg]).to(dev)
                tp, tv = (t.to(dev) for t in pad_targets(tg, device="cpu"))
                tr._graph = None
                mode = "graph"
                try:
                    tr.capture(clips, (tp, tv), sc)
                except Exception:  # noqa: BLE001   (B*priors > 4096: the loss takes the torch formulation, which synchronises)
                    tr._graph = None
                    mode = "eager"
                    torch.cuda.synchronize()
                for _ in range(2):
                    tr.step(clips, (tp, tv), sc)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    tr.step(clips, (tp, tv), sc)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # the slowest rank's time
                ms = float(t)
                tf = B * flop_clip / (ms * 1e-3) / 1e12
                hbm = B_.hbm_estimate(dict(mode="train", frames=T), B, ms, peaks)
                say(f"{T:6d} {B:5d} {ms:8.2f} {world * B * 1000.0 / ms:8.1f} {tf:10.1f} {tf / peak:8.3f} {hbm['achieved']:9.0f} {hbm['frac']:6.3f} "
                    f"{torch.cuda.max_memory_allocated() / 2**30:8.1f}  {mode}")
            except Exception as ex:  # noqa: BLE001
                print(f"[rank {rank}] {T:6d} {B:5d} failed: {repr(ex)[:200]}", flush=True)
            finally:
                tr._graph = None
                tr._graph_out = None
                tr._static = None
                tr._graph_cache.clear()
                torch.cuda.empty_cache()
        del tr, net, crit
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
