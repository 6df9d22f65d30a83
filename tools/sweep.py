"""Clip-length x batch sweep (BASELINE.json configs[4]): training clips/s of one GPU for T in {128,256,512,1024} and
B in {1,2,4,8,16}, with the per-point tensor-pipe roofline fraction (algorithmic conv FLOPs / measured bf16 peak).

    python tools/sweep.py [--frames 128 256 512 1024] [--batches 1 2 4 8 16] > gpurun_out/sweep.txt
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py [--anet] ...     # data parallel, weak scaling

Under torchrun every rank runs the sweep on its own GPU with the gradient all-reduce of `engine.Trainer`; a point's time is
the MAX over ranks (barrier + synchronize on both sides), clips/s is the whole job's, rank 0 prints.

Conv FLOPs scale linearly in T: backbone 0.638 GFLOP/frame forward; the head is sized for T/4 positions.
Each point (bench.measure_point, the same code as the `other_configs` legs of the default bench line): CUDA-graph captured step
on uint8 clips resident in HBM, 2 warm-up + 5 timed replays, CUDA events."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from opental_b200 import engine  # noqa: E402


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
        for B in args.batches:
            try:
                r = B_.measure_point(tr, anet=args.anet, T=T, B=B, dev=dev, world=world, rank=rank, peaks=peaks)
                say(f"{T:6d} {B:5d} {r['ms_per_step']:8.2f} {r['clips_per_s']:8.1f} {r['tensor_tflops_per_gpu']:10.1f} {r['tensor_frac']:8.3f} "
                    f"{r['hbm_gbs']:9.0f} {r['hbm_frac']:6.3f} {r['peak_mem_gb']:8.1f}  {r['mode']}")
            except Exception as ex:  # noqa: BLE001
                print(f"[rank {rank}] {T:6d} {B:5d} failed: {repr(ex)[:200]}", flush=True)
            finally:
                torch.cuda.empty_cache()
        del tr, net, crit
        torch.cuda.empty_cache()
    B_.shutdown()


if __name__ == "__main__":
    main()
