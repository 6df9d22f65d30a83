"""Developer GPU check (not a pytest): runs kernel cases in subprocesses so that a trap in one case does not
poison the others.  Usage on the GPU box:  python tools/gpu_check.py [case ...]  -> gpurun_out/gpu_check.log"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_bmp(inp, seg):
    import torch
    B, C, T = inp.shape
    K = seg.shape[1]
    out = torch.empty(B, C, K, dtype=inp.dtype)
    arg = torch.empty(B, C, K, dtype=torch.long)
    s = seg.to(torch.int64).clamp(0, T - 1)  # trunc toward zero then clamp
    for n in range(B):
        for k in range(K):
            for st in range(2):
                l, r = int(s[n, k, 2 * st]), int(s[n, k, 2 * st + 1])
                cs = slice(st * (C // 2), (st + 1) * (C // 2))
                if r < l:
                    r = l
                win = inp[n, cs, l:r + 1]
                m, a = win.max(dim=1)
                out[n, cs, k] = m
                arg[n, cs, k] = a + l
    return out, arg


def case_bmp():
    import torch
    from opental_b200 import ops
    torch.manual_seed(0)
    res = {}
    for (B, C, T, K) in [(2, 64, 32, 32), (1, 1024, 64, 64), (3, 512, 256, 16), (2, 1024, 2, 2), (1, 512, 256, 64)]:
        inp = torch.randn(B, C, T)
        c = torch.rand(B, K, 1) * T
        seg = torch.cat([c - torch.rand(B, K, 1) * 12 - 2, c + torch.rand(B, K, 1) * 6,
                         c - torch.rand(B, K, 1) * 6, c + torch.rand(B, K, 1) * 12 + 2], -1).round()
        ro, ra = ref_bmp(inp, seg)
        out = ops.bmp_forward(inp.cuda(), seg.cuda()).cpu()
        go = torch.randn(B, C, K)
        gi_ref = torch.zeros(B, C, T).scatter_add_(2, ra, go)
        gi = ops.bmp_backward(go.cuda(), inp.cuda(), seg.cuda(), False).cpu()
        res[f"{B}x{C}x{T}x{K}"] = dict(fwd_equal=bool(torch.equal(out, ro)), bwd_maxabs=float((gi - gi_ref).abs().max()))
    return res


def conv_case(N, T, H, W, Cin, Cout, k, nsplit, relu=True, seed=0):
    import torch
    import torch.nn.functional as F
    from opental_b200 import ops
    torch.manual_seed(seed)
    dev = "cuda"
    x = torch.randn(N, Cin, T, H, W, device=dev)
    w = torch.randn(Cout, Cin, *k, device=dev) * (2.0 / (Cin * k[0] * k[1] * k[2])) ** 0.5
    scale = 1 + 0.1 * torch.randn(Cout, device=dev)
    shift = 0.1 * torch.randn(Cout, device=dev)
    pads = [(kk - 1) // 2 for kk in k]
    pad_full = []
    for kk in reversed(k):
        pad_full += [(kk - 1) // 2, kk - 1 - (kk - 1) // 2]
    ref = F.conv3d(F.pad(x.double(), pad_full), w.double()) * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1)
    if relu:
        ref = ref.relu()
    xp = ops.split_bf16(x.permute(0, 2, 3, 4, 1).contiguous(), with_lo=nsplit == 3)
    wp = ops.pack_conv_weight(w, with_lo=nsplit == 3)
    of32 = torch.zeros(N, T, H, W, Cout, device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    out = ops.conv_igemm(xp, wp, kernel=tuple(k), pad_front=tuple(pads), scale=scale, shift=shift, relu=relu, out_f32=of32)
    torch.cuda.synchronize()
    dt = time.time() - t0
    refl = ref.permute(0, 2, 3, 4, 1)
    y = out.float().double()
    e_planes = float((y - refl).abs().max() / refl.abs().max())
    e_f32 = float((of32.double() - refl).abs().max() / refl.abs().max())
    rms = float(((of32.double() - refl).pow(2).mean() / refl.pow(2).mean()).sqrt())
    # timing
    for _ in range(2):
        ops.conv_igemm(xp, wp, kernel=tuple(k), pad_front=tuple(pads), scale=scale, shift=shift, relu=relu, out=out)
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    reps = 5
    for _ in range(reps):
        ops.conv_igemm(xp, wp, kernel=tuple(k), pad_front=tuple(pads), scale=scale, shift=shift, relu=relu, out=out)
    en.record()
    torch.cuda.synchronize()
    ms = st.elapsed_time(en) / reps
    flops = 2.0 * N * T * H * W * Cout * Cin * k[0] * k[1] * k[2]
    return dict(err_planes=e_planes, err_f32=e_f32, rel_rms=rms, first_call_s=dt, ms=ms, tflops_alg=flops / ms / 1e9)


CONV_CASES = {
    "conv_1x1_small": dict(N=1, T=4, H=8, W=8, Cin=64, Cout=64, k=(1, 1, 1), nsplit=1),
    "conv_1x1_x3": dict(N=1, T=4, H=8, W=8, Cin=64, Cout=64, k=(1, 1, 1), nsplit=3),
    "conv_3x3_small": dict(N=2, T=8, H=12, W=12, Cin=96, Cout=208, k=(3, 3, 3), nsplit=3),
    "conv_3x3_odd": dict(N=1, T=5, H=6, W=6, Cin=24, Cout=64, k=(3, 3, 3), nsplit=3),
    "conv_1d": dict(N=2, T=64, H=1, W=1, Cin=512, Cout=512, k=(3, 1, 1), nsplit=3),
    "conv_nblocks": dict(N=1, T=16, H=3, W=3, Cin=832, Cout=384, k=(1, 1, 1), nsplit=3),
    "conv_2c": dict(N=1, T=128, H=24, W=24, Cin=64, Cout=192, k=(3, 3, 3), nsplit=3),
    "conv_2c_bf16": dict(N=1, T=128, H=24, W=24, Cin=64, Cout=192, k=(3, 3, 3), nsplit=1),
    "conv_3c_b1b": dict(N=1, T=128, H=12, W=12, Cin=128, Cout=192, k=(3, 3, 3), nsplit=3),
}


def run_case(name):
    if name == "bmp":
        return case_bmp()
    return conv_case(**CONV_CASES[name])


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        try:
            print("RESULT " + json.dumps(run_case(sys.argv[2])))
        except Exception as e:  # noqa: BLE001
            print("RESULT " + json.dumps({"error": repr(e)[:800]}))
        sys.exit(0)
    names = sys.argv[1:] or ["bmp", *CONV_CASES]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "a")
    for n in names:
        try:
            pr = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=300)
            lines = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            msg = lines[-1][7:] if lines else json.dumps({"rc": pr.returncode, "stdout": pr.stdout[-1500:], "stderr": pr.stderr[-1500:]})
        except subprocess.TimeoutExpired:
            msg = json.dumps({"error": "timeout"})
        line = f"{n}: {msg}"
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()
