"""Developer tool: per-tile pipeline timeline of conv_igemm_kernel (needs the library built with OTAL_BUILD_DEFINES=OTAL_TIMELINE).
For a few representative launches prints, per role, where the clock64() time of a tile goes — who waits for whom.
    OTAL_BUILD_DEFINES=OTAL_TIMELINE python -m opental_b200.build && gpurun -- python tools/conv_timeline.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from opental_b200 import _lib, ops  # noqa: E402

SLOTS = ["prod_start", "prod_first", "prod_last", "mma_first_full", "mma_commit", "epi_wait", "epi_got", "epi_chunk0", "epi_release",
         "mma_acc_wait", "mma_acc_got"]


def run(label, fn, grid=148):
    lib = _lib.load()
    buf = torch.zeros(grid * 16 * 16, dtype=torch.int64, device="cuda")
    lib.otal_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
    fn(); torch.cuda.synchronize()                      # warm (tensor maps, L2 state comparable to a steady step)
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.otal_debug_set_timeline(ctypes.c_void_p(0))
    t = buf.view(grid, 16, 16).cpu()
    print(f"\n=== {label}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    for cta in (0, 77):
        base = int(t[cta, 0, 0])
        print(f" CTA {cta}: per tile, cycles since the CTA's first producer event")
        print("  tile " + " ".join(f"{s[:11]:>11s}" for s in SLOTS))
        for tile in range(12):
            if int(t[cta, tile, 0]) == 0:
                break
            print(f"  {tile:4d} " + " ".join(f"{int(t[cta, tile, k]) - base:11d}" if int(t[cta, tile, k]) else f"{'-':>11s}" for k in range(len(SLOTS))))
    # averages over all CTAs: tile period and the main waits
    v = t[:, :12, :].double()
    ok = v[:, :, 0] > 0
    per = (v[:, 1:, 8] - v[:, :-1, 8])[ok[:, 1:] & (v[:, 1:, 8] > 0)]
    print(f" mean tile period (epilogue release to release): {float(per.mean()):.0f} cycles; "
          f"producer start->last issue {float((v[:, :, 2] - v[:, :, 0])[ok].mean()):.0f}; "
          f"MMA first full->commit {float((v[:, :, 4] - v[:, :, 3])[ok].mean()):.0f}; MMA wait for acc {float((v[:, :, 10] - v[:, :, 9])[ok].mean()):.0f}; "
          f"epilogue wait {float((v[:, :, 6] - v[:, :, 5])[ok].mean()):.0f}, got->chunk0 {float((v[:, :, 7] - v[:, :, 6])[ok].mean()):.0f}, "
          f"got->release {float((v[:, :, 8] - v[:, :, 6])[ok].mean()):.0f}")


def main():
    torch.manual_seed(0)
    dev = "cuda"
    N = 8

    def planes(*shape):
        return ops.split_bf16(torch.randn(*shape, device=dev))

    # 1. Mixed_3c.b0 forward: 1x1, 256 -> 128 into a 480-wide concat buffer
    x = planes(N, 128, 12, 12, 256); w = ops.pack_conv_weight(torch.randn(128, 256, 1, 1, 1, device=dev) * 0.05)
    y = planes(N, 128, 12, 12, 480); sc = torch.ones(128, device=dev); sh = torch.zeros(128, device=dev)
    run("Mixed_3c.b0 fwd 1x1 256->128 @8x128x12x12 (planes out)",
        lambda: ops.conv_igemm(x, w, kernel=(1, 1, 1), pad_front=(0, 0, 0), scale=sc, shift=sh, relu=True, out=y, out_slice=(0, 128)))
    # 2. Mixed_3c.b1b forward: 3x3x3, 128 -> 192
    m = planes(N, 128, 12, 12, 160); w3 = ops.pack_conv_weight(torch.randn(192, 128, 3, 3, 3, device=dev) * 0.02)
    sc3 = torch.ones(192, device=dev); sh3 = torch.zeros(192, device=dev)
    run("Mixed_3c.b1b fwd 3x3x3 128->192 @8x128x12x12",
        lambda: ops.conv_igemm(m, w3, kernel=(3, 3, 3), pad_front=(1, 1, 1), scale=sc3, shift=sh3, relu=True, in_slice=(0, 128), out=y,
                               out_slice=(128, 192)))
    # 3. Conv3d_2b data gradient: 1x1, d [.,64] -> fp32 g [.,64]
    d = planes(N, 128, 24, 24, 64); w2 = ops.pack_conv_weight(torch.randn(64, 64, 1, 1, 1, device=dev) * 0.1)
    g = torch.empty(N, 128, 24, 24, 64, device=dev)
    run("Conv3d_2b dgrad 1x1 64->64 @8x128x24x24 (fp32 out)",
        lambda: ops.conv_igemm(d, w2, kernel=(1, 1, 1), pad_front=(0, 0, 0), out_f32=g, want_planes=False, dgrad=True))
    # 4. Conv3d_2b forward (planes out)
    y2 = planes(N, 128, 24, 24, 64); s64 = torch.ones(64, device=dev); z64 = torch.zeros(64, device=dev)
    run("Conv3d_2b fwd 1x1 64->64 @8x128x24x24 (planes out)",
        lambda: ops.conv_igemm(d, w2, kernel=(1, 1, 1), pad_front=(0, 0, 0), scale=s64, shift=z64, relu=True, out=y2))
    # 5. Mixed_4c.b0-like: 512 -> 160 at 8x64x6x6
    x4 = planes(N, 64, 6, 6, 512); w4 = ops.pack_conv_weight(torch.randn(160, 512, 1, 1, 1, device=dev) * 0.05)
    y4 = planes(N, 64, 6, 6, 512); s160 = torch.ones(160, device=dev); z160 = torch.zeros(160, device=dev)
    run("Mixed_4c.b0 fwd 1x1 512->160 @8x64x6x6",
        lambda: ops.conv_igemm(x4, w4, kernel=(1, 1, 1), pad_front=(0, 0, 0), scale=s160, shift=z160, relu=True, out=y4, out_slice=(0, 160)))


if __name__ == "__main__":
    main()
