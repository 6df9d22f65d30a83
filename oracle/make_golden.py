"""Generate tests/golden/* from the reference's own code (build container only)  —  TEST INFRASTRUCTURE.

    python oracle/make_golden.py

1. pins the oracle restatement (oracle/opental_oracle.py) against the reference imported from /root/reference
   (asserts agreement, see TOL below), and
2. writes small fixtures (inputs are re-generated from seeds, only outputs are stored) that the CPU tests
   re-check the oracle against and the GPU tests check the CUDA path against.
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import opental_oracle as O  # noqa: E402
import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5  # max |ref - oracle| / max |ref| ; both are fp32 CPU torch, differences are summation-order noise


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def bmp_cases():
    """Small pooling cases from the scalar kernel emulation (incl. T != K compat backward, ragged windows)."""
    g = torch.Generator().manual_seed(1234)
    cases = {}
    for name, (B, C, T, K) in dict(level=(2, 8, 16, 16), frame=(2, 6, 32, 8), ssl=(1, 4, 64, 3), tiny=(1, 2, 2, 2),
                                   wide=(1, 4, 40, 5)).items():
        inp = torch.randn(B, C, T, generator=g)
        c = torch.rand(B, K, 1, generator=g) * T
        seg = torch.cat([c - torch.rand(B, K, 1, generator=g) * 9 - 1, c + torch.rand(B, K, 1, generator=g) * 5,
                         c - torch.rand(B, K, 1, generator=g) * 5, c + torch.rand(B, K, 1, generator=g) * 9 + 1], -1)
        if name == "wide":
            seg = seg / 1.7          # fractional -> truncation matters; includes negatives and > T-1
            seg[0, 0] = torch.tensor([5.0, 2.0, -3.0, 100.0])  # r < l window, clamping both sides
        else:
            seg = seg.round()
        inp[0, 0, : min(4, T)] = 1.5  # ties: first index must win
        gout = torch.randn(B, C, K, generator=g)
        fwd = ref_loader.kernel_emulation_forward(inp, seg)
        bwd_compat = ref_loader.kernel_emulation_backward(gout, inp, seg, compat=True) if K <= T else None
        bwd_fixed = ref_loader.kernel_emulation_backward(gout, inp, seg, compat=False)
        # pin the vectorised oracle against the scalar emulation
        x = inp.clone().requires_grad_(True)
        y = O.boundary_max_pooling(x, seg, True)
        assert torch.equal(y.detach(), fwd), name
        if bwd_compat is not None:
            (gx,) = torch.autograd.grad(y, x, gout)
            assert torch.allclose(gx, bwd_compat, atol=1e-6), name
        x = inp.clone().requires_grad_(True)
        (gx,) = torch.autograd.grad(O.boundary_max_pooling(x, seg, False), x, gout)
        assert torch.allclose(gx, bwd_fixed, atol=1e-6), name
        cases[name] = dict(inp=inp.numpy(), seg=seg.numpy(), gout=gout.numpy(), fwd=fwd.numpy(),
                           bwd_fixed=bwd_fixed.numpy(),
                           **({"bwd_compat": bwd_compat.numpy()} if bwd_compat is not None else {}))
    flat = {f"{n}.{k}": v for n, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(GOLD, "bmp_cases.npz"), **flat)
    print("bmp cases:", list(cases))


def model_cases():
    ns = ref_loader.load_reference()
    cfg = O.OracleConfig()
    summary = {}
    arrays = {}
    for tag, shift in (("init", 0.0), ("biased", math.log(32.0))):
        sd = O.synthetic_state_dict(cfg, loc_bias_shift=shift)
        net = ns.BDNet(in_channels=3, training=False, use_edl=True)
        net.load_state_dict(sd)
        net.train()  # BN stays frozen (BDNet.py:39-49); dropout p=0
        x = O.synthetic_clip(0).unsqueeze(0)
        targets = [O.synthetic_targets(0, num_classes=cfg.num_classes)]
        scores = O.synthetic_scores(targets[0]).unsqueeze(0)
        for epoch in (1, 11):
            crit = ns.MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=ns.config["training"]["edl_config"],
                                       os_head=True, act_config=ns.config["training"]["act_config"])
            crit.cls_loss.epoch = epoch
            net.zero_grad()
            out_r = net(x)
            loss_r = crit(out_r, [t.clone() for t in targets])
            cost_r = loss_r[0] + 10 * loss_r[1] + loss_r[2] + 10 * loss_r[3] + loss_r[4] + loss_r[5] + loss_r[6]
            cost_r.backward()
            grads_r = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}

            sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
            state = O.LossState(epoch=epoch)
            out_o = O.bdnet_forward(x, sdo, cfg, compat=True)
            loss_o = O.multisegment_loss(out_o, targets, state, cfg)
            cost_o = loss_o[0] + 10 * loss_o[1] + loss_o[2] + 10 * loss_o[3] + loss_o[4] + loss_o[5] + loss_o[6]
            cost_o.backward()

            errs = {k: rel(out_o[k].detach(), out_r[k].detach()) for k in out_r if out_r[k] is not None}
            lerrs = [abs(float(a) - float(b)) / max(abs(float(b)), 1e-6) for a, b in zip(loss_o, loss_r)]
            gerrs = {k: rel(sdo[k].grad, g) for k, g in grads_r.items() if sdo[k].grad is not None}
            missing = [k for k in grads_r if sdo[k].grad is None and grads_r[k].abs().max() > 0]
            worst_out, worst_g = max(errs.values()), max(gerrs.values())
            print(f"[{tag} epoch {epoch}] oracle vs reference: outputs {worst_out:.2e} losses {max(lerrs):.2e} grads {worst_g:.2e}"
                  f" (n={len(gerrs)}) missing={missing}")
            assert worst_out < TOL and max(lerrs) < 5e-5 and not missing, (errs, lerrs)
            # Gradients are discontinuous in the activations (ReLU masks, max-pool / boundary-pool argmax): a
            # 1e-6 forward difference flips a handful of mask bits and moves a weight gradient by ~1e-2 relative
            # (measured: the reference differs from ITSELF by ~1e-3 between 3 and 8 CPU threads, and from an fp64
            # run by 3-9e-3, while this restatement matches fp64 to 1e-6 on the late layers).  Hence the bound.
            assert worst_g < 5e-2, sorted(gerrs.items(), key=lambda kv: -kv[1])[:5]
            w_acc_ref = crit.cls_loss.weight_accum.clone()
            assert torch.allclose(w_acc_ref, state.weight_accum, atol=1e-6)

            key = f"{tag}.e{epoch}"
            summary[key] = dict(losses=[float(v) for v in loss_r], cost=float(cost_r),
                                n_pos=int((loss_r[0] > 0)), oracle_vs_ref_out=worst_out, oracle_vs_ref_grad=worst_g)
            if epoch == 1:
                for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "unct", "prop_unct"):
                    arrays[f"{tag}.{k}"] = out_r[k].detach().numpy()
                for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
                    arrays[f"{tag}.{k}.sample"] = out_r[k].detach().numpy()[:, ::8, ::8].copy()
                    arrays[f"{tag}.{k}.mean"] = np.array([out_r[k].mean().item(), out_r[k].abs().mean().item()])
            arrays[f"{key}.weight_accum"] = w_acc_ref.numpy()
            # gradient fingerprints: (sum, abs-sum, strided sample) for every trainable tensor
            fp = {}
            for k, g in grads_r.items():
                fp[k] = [float(g.sum()), float(g.abs().sum())]
                arrays[f"{key}.grad.{k}"] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy().copy()
            summary[key]["grad_fingerprint"] = fp
        # backbone endpoints (strided samples) for the layer-wise conv parity tests
        with torch.no_grad():
            feats = O.i3d_features(x, sd, keep=None)
        for name, f in feats.items():
            arrays[f"{tag}.feat.{name}.sample"] = f[0, ::7, ::5, ::3, ::3].numpy().copy()
            arrays[f"{tag}.feat.{name}.stats"] = np.array([f.mean().item(), f.abs().mean().item(), f.max().item()])
        if tag == "init":
            # and the reference's own endpoints agree with the oracle's
            with torch.no_grad():
                fr = net.backbone(x)
            for name in feats:
                assert rel(feats[name], fr[name]) < TOL, name
    np.savez_compressed(os.path.join(GOLD, "model_thumos_opental.npz"), **arrays)
    with open(os.path.join(GOLD, "model_thumos_opental.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    ns.restore_cuda()


def anet_cases():
    """ActivityNet flavour (configs/anet_opental.yaml --open_set): 768-frame clip, 150 classes, 189 priors, the
    per-sample loss.  Run in its own process (`--anet`): the reference's config is a per-process global."""
    ns = ref_loader.load_reference(config="configs/anet_opental.yaml", extra_args=("--open_set", "--split=0"), flavour="anet")
    cfg = O.anet_config()
    summary, arrays = {}, {}
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=math.log(8.0))      # x fpn stride 4..128 -> extents of 32..1024 frames
    net = ns.BDNet(in_channels=3, training=False, frame_num=768, use_edl=True)
    net.load_state_dict(sd)
    net.train()
    B = 2
    x = torch.stack([O.synthetic_clip(i, frames=768) for i in range(B)])
    targets = [O.synthetic_targets(i, num_classes=cfg.num_classes) for i in range(B)]
    targets[1] = torch.cat([targets[1], torch.tensor([[0.40, 0.44, 17.0]])])          # a short third segment (level gating)
    for epoch in (1, 11):
        crit = ns.MultiSegmentLoss(cfg.num_classes, 0.5, 1.0, cls_loss_type="edl", edl_config=ns.config["training"]["edl_config"],
                                   os_head=True)
        crit.cls_loss.epoch = epoch
        net.zero_grad()
        out_r = net(x)
        loss_r = crit([out_r[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "priors", "act", "prop_act")],
                      [t.clone() for t in targets])
        cost_r = loss_r[0] + 10 * loss_r[1] + loss_r[2] + 10 * loss_r[3] + loss_r[4] + loss_r[5] + loss_r[6]
        cost_r.backward()
        grads_r = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}

        sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        state = O.LossState(epoch=epoch)
        out_o = O.bdnet_forward(x, sdo, cfg, compat=True)
        loss_o = O.multisegment_loss_anet(out_o, targets, state, cfg)
        cost_o = loss_o[0] + 10 * loss_o[1] + loss_o[2] + 10 * loss_o[3] + loss_o[4] + loss_o[5] + loss_o[6]
        cost_o.backward()
        errs = {k: rel(out_o[k].detach(), out_r[k].detach()) for k in out_r if out_r[k] is not None}
        lerrs = [abs(float(a) - float(b)) / max(abs(float(b)), 1e-6) for a, b in zip(loss_o, loss_r)]
        gerrs = {k: rel(sdo[k].grad, g) for k, g in grads_r.items() if sdo[k].grad is not None}
        missing = [k for k in grads_r if sdo[k].grad is None and grads_r[k].abs().max() > 0]
        worst_out, worst_g = max(errs.values()), max(gerrs.values())
        print(f"[anet epoch {epoch}] oracle vs reference: outputs {worst_out:.2e} losses {max(lerrs):.2e} grads {worst_g:.2e}"
              f" (n={len(gerrs)}) missing={missing}  losses={[round(float(v), 4) for v in loss_r]}")
        assert worst_out < TOL and max(lerrs) < 5e-5 and not missing, (errs, lerrs)
        assert worst_g < 5e-2, sorted(gerrs.items(), key=lambda kv: -kv[1])[:5]
        key = f"anet.e{epoch}"
        summary[key] = dict(losses=[float(v) for v in loss_r], cost=float(cost_r), oracle_vs_ref_out=worst_out,
                            oracle_vs_ref_grad=worst_g)
        if epoch == 1:
            for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "unct", "prop_unct", "priors"):
                arrays[f"anet.{k}"] = out_r[k].detach().numpy()
            for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
                arrays[f"anet.{k}.sample"] = out_r[k].detach().numpy()[:, ::8, ::8].copy()
        fp = {}
        for k, g in grads_r.items():
            fp[k] = [float(g.sum()), float(g.abs().sum())]
            arrays[f"{key}.grad.{k}"] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy().copy()
        summary[key]["grad_fingerprint"] = fp
    np.savez_compressed(os.path.join(GOLD, "model_anet_opental.npz"), **arrays)
    with open(os.path.join(GOLD, "model_anet_opental.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    ns.restore_cuda()


def ssl_cases():
    """The SSL / triplet second pass of the THUMOS14 training step (train.py:174-184, 237-242; BDNet.py:482-503)."""
    import torch.nn as nn
    ns = ref_loader.load_reference()
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=math.log(32.0))
    net = ns.BDNet(in_channels=3, training=False, use_edl=True)
    net.load_state_dict(sd)
    net.train()
    x = O.synthetic_clip(1).unsqueeze(0)
    proposals = [torch.tensor([[40.0, 90.0], [122.0, 172.0], [91.0, 121.0]])]      # anchor / positive / negative, frames
    net.zero_grad()
    a_r, p_r, n_r = net(x, proposals=proposals, ssl=True)
    trip_r = torch.stack([nn.TripletMarginLoss()(a_r[i], p_r[i], n_r[i]) * w for i, w in enumerate((1, 0.1, 0.1))]).sum(0)
    trip_r.backward()
    grads_r = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    a_o, p_o, n_o = O.bdnet_forward_ssl(x, sdo, cfg, proposals, compat=True)
    trip_o = O.triplet_loss(a_o, p_o, n_o)
    trip_o.backward()
    errs = [rel(o.detach(), r.detach()) for lo, lr in ((a_o, a_r), (p_o, p_r), (n_o, n_r)) for o, r in zip(lo, lr)]
    gerrs = {k: rel(sdo[k].grad, g) for k, g in grads_r.items() if sdo[k].grad is not None and g.abs().max() > 0}
    print(f"[ssl] oracle vs reference: features {max(errs):.2e} triplet {abs(float(trip_o) - float(trip_r)):.2e} "
          f"grads {max(gerrs.values()):.2e} (n={len(gerrs)})  trip={float(trip_r):.5f}")
    assert max(errs) < TOL and abs(float(trip_o) - float(trip_r)) < 1e-5 * max(1.0, abs(float(trip_r)))
    assert max(gerrs.values()) < 5e-2
    arrays = {}
    for name, lst in (("anchor", a_r), ("positive", p_r), ("negative", n_r)):
        for i, t in enumerate(lst):
            arrays[f"ssl.{name}.{i}"] = t.detach().numpy()
    fp = {}
    for k, g in grads_r.items():
        fp[k] = [float(g.sum()), float(g.abs().sum())]
        arrays[f"ssl.grad.{k}"] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "model_thumos_ssl.npz"), **arrays)
    with open(os.path.join(GOLD, "model_thumos_ssl.json"), "w") as fh:
        json.dump(dict(triplet=float(trip_r), proposals=proposals[0].tolist(), grad_fingerprint=fp), fh, indent=1)
    ns.restore_cuda()


def infer_cases():
    """Inference post-processing (test.py:112-162, segment_utils.py:128-162): reference functions vs the oracle."""
    import importlib
    ns = ref_loader.load_reference()
    ref_test = importlib.import_module("AFSD.thumos14.test")
    ref_seg = importlib.import_module("AFSD.common.segment_utils")
    cfg = O.OracleConfig()
    g = torch.Generator().manual_seed(77)
    P, K = 126, cfg.num_classes
    out = dict(loc=torch.rand(1, P, 2, generator=g) * 40 + 1, conf=2 * torch.randn(1, P, K, generator=g),
               prop_loc=0.3 * torch.randn(1, P, 2, generator=g), prop_conf=2 * torch.randn(1, P, K, generator=g),
               center=torch.randn(1, P, 1, generator=g), priors=torch.cat(O.level_priors(cfg), 0),
               act=2 * torch.randn(1, P, 1, generator=g), prop_act=2 * torch.randn(1, P, 1, generator=g))
    out["unct"], out["prop_unct"] = O.dirichlet_uncertainty(out["conf"]), O.dirichlet_uncertainty(out["prop_conf"])
    layer = ns.bdnet_module.DirichletLayer(evidence="exp", dim=-1)
    loc, conf, ploc, pconf, center, priors, unct, punct, act, pact = ref_test.parse_output(out, use_edl=True, os_head=True)
    seg_r, sc_r, un_r, ac_r = ref_test.decode_predictions(loc, ploc, priors, conf, pconf, unct, punct, act, pact, center, 384, 10.0,
                                                          256, K, score_func=layer, use_edl=True, os_head=True)
    seg_o, sc_o, un_o, ac_o = O.decode_predictions(out, 0, 384, 10.0, cfg)
    for a, b in ((seg_o, seg_r), (sc_o, sc_r), (un_o, un_r), (ac_o, ac_r)):
        assert rel(a, b) < 1e-6
    arrays = dict(seg=seg_r.numpy(), scores=sc_r.numpy(), unct=un_r.numpy(), act=ac_r.numpy())
    n_f = 0
    for cl in range(K):
        fr = ref_test.filtering(seg_r, sc_r[cl], un_r, ac_r, 0.001, use_edl=True, os_head=True)
        fo = O.filter_candidates(seg_o, sc_o[cl], un_o, ac_o, 0.001)
        assert (fr is None) == (fo is None)
        if fr is not None:
            assert torch.allclose(fr, fo, atol=1e-6)
            n_f += fr.shape[0]
    # soft-NMS cases: random candidates, a case with duplicates / ties, and one that hits top_k
    for name, (n, top_k, sigma) in dict(a=(300, 1000, 0.5), b=(64, 10, 0.85), c=(5, 1000, 0.5), d=(1, 1000, 0.5)).items():
        st = torch.rand(n, generator=g) * 100
        cand = torch.stack([st, st + torch.rand(n, generator=g) * 30 + 0.5, torch.rand(n, generator=g), torch.rand(n, generator=g),
                            torch.rand(n, generator=g)], -1)
        if name == "b":
            cand[10] = cand[3]                         # exact duplicate: first index wins the argmax
        kept_r, cnt_r, mask_r = ref_seg.softnms_v2(cand.clone(), sigma=sigma, top_k=top_k, use_edl=True, os_head=True, get_mask=True)
        kept_o, cnt_o, mask_o = O.softnms_v2(cand.clone(), sigma=sigma, top_k=top_k)
        assert int(cnt_r) == cnt_o and torch.equal(mask_r, mask_o) and torch.allclose(kept_r, kept_o, atol=1e-7), name
        arrays[f"nms.{name}.cand"] = cand.numpy()
        arrays[f"nms.{name}.kept"] = kept_r.numpy()
        arrays[f"nms.{name}.mask"] = mask_r.numpy()
        arrays[f"nms.{name}.cfg"] = np.array([top_k, sigma])
    print(f"[infer] decode / filtering ({n_f} candidates) / soft-NMS: oracle == reference")
    np.savez_compressed(os.path.join(GOLD, "infer_cases.npz"), **arrays)
    ns.restore_cuda()


def edl_cases():
    """Every branch of EvidenceLoss (cls_loss.py:81-285) on seeded logits, called repeatedly so the stateful branches
    (GHM acc_sum, IBM weight_accum) evolve: reference vs oracle, loss value and gradient."""
    import importlib
    ns = ref_loader.load_reference()
    ref_cls = importlib.import_module("AFSD.thumos14.cls_loss")
    K = 15
    arrays = {}
    for name, (cfg, epochs) in O.EDL_VARIANTS.items():
        size_average = name.endswith("_mean")
        ref = ref_cls.EvidenceLoss(K, dict(cfg), size_average=size_average)
        st = O.EdlVariantState(cfg.get("num_bins", 50))
        for call, epoch in enumerate(epochs):
            logit, target = O.edl_inputs(name, call, K)
            ref.epoch = st.epoch = epoch
            lr = logit.clone().requires_grad_(True)
            loss_r = ref(lr, target)
            loss_r.backward()
            lo = logit.clone().requires_grad_(True)
            loss_o = O.evidence_loss_variant(lo, target, st, K, cfg, size_average=size_average)
            loss_o.backward()
            assert abs(float(loss_r) - float(loss_o)) <= TOL * max(1.0, abs(float(loss_r))), (name, call, float(loss_r), float(loss_o))
            assert rel(lo.grad, lr.grad) < TOL, (name, call, rel(lo.grad, lr.grad))
            arrays[f"{name}.{call}.loss"] = np.array(float(loss_r), dtype=np.float64)
            arrays[f"{name}.{call}.grad"] = lr.grad.numpy().copy()
        if cfg.get("with_ibm"):
            assert torch.allclose(ref.weight_accum, st.weight_accum, atol=1e-7)
            arrays[f"{name}.weight_accum"] = ref.weight_accum.numpy().copy()
        if cfg.get("with_ghm") and cfg.get("momentum", 0) > 0:
            assert np.allclose(ref.acc_sum, st.acc_sum, rtol=1e-12)
            arrays[f"{name}.acc_sum"] = np.array(ref.acc_sum)
        # the same configuration inside the whole MultiSegmentLoss (reference values only: the product's masked
        # formulation is checked against them in tests/test_loss_cpu.py)
        msl = ns.MultiSegmentLoss(K, 0.5, 1.0, use_gpu=False, cls_loss_type="edl", edl_config=dict(cfg, iou_aware=True), os_head=True,
                                  act_config=dict(weight=0.1, margin=1.0))
        msl.cls_loss.epoch = 11
        for it in range(2):
            out = {k: v.requires_grad_(True) for k, v in O.fake_head_outputs(3, 5 + it, K).items()}
            out["priors"] = torch.cat(O.level_priors(O.OracleConfig()), 0)
            losses = msl(out, [O.synthetic_targets(i, num_classes=K) for i in range(3)])
            keys = ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act")
            grads = torch.autograd.grad(sum(w * l for w, l in zip((1, 10, 1, 10, 1, 1, 1), losses)), [out[k] for k in keys])
            arrays[f"{name}.msl.{it}.losses"] = np.array([float(l) for l in losses])
            for k, gk in zip(keys, grads):
                arrays[f"{name}.msl.{it}.grad.{k}"] = gk.numpy().copy()
        print(f"[edl] {name}: oracle == reference over {len(epochs)} calls")
    np.savez_compressed(os.path.join(GOLD, "edl_variants.npz"), **arrays)
    ns.restore_cuda()


def augment_case_inputs(seed: int):
    """Seeded annotation sets (frames) + threshold for the cut-paste fixtures."""
    import random
    r = random.Random(1000 + seed)
    n = r.choice([1, 2, 2, 3])
    cuts = sorted(r.sample(range(2, 254), 2 * n))
    annos = [[cuts[2 * i], cuts[2 * i + 1], r.randint(1, 15)] for i in range(n)]
    if seed % 5 == 0:
        annos = [[a[0] + 0.5, a[1] + 0.25, a[2]] for a in annos]       # fractional boundaries (ceil / floor paths)
    return annos, r.choice([3, 5, 8, 13, 21])


def augment_cases():
    """SSL cut-paste augmentation (thumos_dataset.py:172-237): the reference's own method run on a clip whose pixel value
    is its frame index, so the augmented clip IS the frame map; same `random` seed on both sides."""
    import importlib
    import random
    import types
    ns = ref_loader.load_reference()
    ds = importlib.import_module("AFSD.common.thumos_dataset")
    from opental_b200 import augment as A
    dummy = types.SimpleNamespace(clip_length=256)
    dummy.get_bg = types.MethodType(ds.THUMOS_Dataset.get_bg, dummy)
    dummy.augment_ = types.MethodType(ds.THUMOS_Dataset.augment_, dummy)
    cases, n_ok = [], 0
    clip = torch.arange(256, dtype=torch.float32).view(1, 256, 1, 1).expand(3, 256, 2, 2).contiguous()
    for seed in range(60):
        annos, th = augment_case_inputs(seed)
        random.seed(seed)
        new_input, new_annos, flag = ds.THUMOS_Dataset.augment(dummy, clip, [list(a) for a in annos], th, 1)
        ref_map = new_input[0, :, 0, 0].long().tolist()
        assert torch.equal(new_input, clip[:, ref_map])                 # it is a pure frame re-ordering
        random.seed(seed)
        fmap, got_annos, got_flag = A.cut_paste([list(a) for a in annos], th, 256, 1)
        assert got_flag == flag and fmap.tolist() == ref_map, seed
        assert [list(map(float, a)) for a in got_annos] == [list(map(float, a)) for a in new_annos], seed
        n_ok += bool(flag)
        cases.append(dict(seed=seed, annos=annos, th=th, flag=bool(flag), frame_map=ref_map,
                          new_annos=[list(map(float, a)) for a in new_annos]))
    print(f"[augment] 60 seeds ({n_ok} augmented): frame map == the reference's augmented clip")
    with open(os.path.join(GOLD, "augment_cases.json"), "w") as fh:
        json.dump(cases, fh)
    ns.restore_cuda()


def closed_cfg():
    return O.OracleConfig(num_classes=21, os_head=False, use_edl=False, with_ibm=False, iou_aware=False)


def config1_clip():
    """x = (randint(0,256,[1,3,256,96,96], seed 0).float()/255)*2-1."""
    g = torch.Generator().manual_seed(0)
    return (torch.randint(0, 256, [1, 3, 256, 96, 96], generator=g).float() / 255) * 2 - 1


def closed_cases():
    """BASELINE configs[0] / SURVEY §8d config 1: `configs/thumos14.yaml` (closed set: 21 classes incl. background, softmax
    focal loss, no EDL, no actionness heads), `BDNet(in_channels=3, training=False).eval()` forward on the seed-0 clip; plus
    one training step's losses and gradient fingerprints with MultiSegmentLoss(cls_loss_type='focal')."""
    ns = ref_loader.load_reference(config="configs/thumos14.yaml", extra_args=())
    cfg = closed_cfg()
    x = config1_clip()
    arrays, summary = {}, {}
    for tag, shift in (("init", 0.0), ("biased", math.log(32.0))):
        sd = O.synthetic_state_dict(cfg, loc_bias_shift=shift)
        net = ns.BDNet(in_channels=3, training=False)
        net.load_state_dict(sd)
        net.eval()
        with torch.no_grad():
            out_r = net(x)
            out_o = O.bdnet_forward(x, sd, cfg, compat=True)
        assert out_r["act"] is None and out_r["prop_act"] is None and "unct" not in out_r
        errs = {k: rel(out_o[k], out_r[k]) for k in out_r if out_r[k] is not None}
        assert max(errs.values()) < TOL, errs
        for k in ("loc", "conf", "prop_loc", "prop_conf", "center"):
            arrays[f"{tag}.{k}"] = out_r[k].numpy()
        for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
            arrays[f"{tag}.{k}.sample"] = out_r[k].numpy()[:, ::8, ::8].copy()
        # training step with the focal loss
        net.train()
        targets = [O.synthetic_targets(0, num_classes=20)]
        crit = ns.MultiSegmentLoss(21, 0.5, 1.0, use_gpu=False, cls_loss_type="focal")
        net.zero_grad()
        out_t = net(x)
        loss_r = crit(out_t, [t.clone() for t in targets])
        assert loss_r[5] is None and loss_r[6] is None
        cost_r = loss_r[0] + 10 * loss_r[1] + loss_r[2] + 10 * loss_r[3] + loss_r[4]
        cost_r.backward()
        sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        out_ot = O.bdnet_forward(x, sdo, cfg, compat=True)
        loss_o = O.multisegment_loss_closed(out_ot, targets, cfg)
        (loss_o[0] + 10 * loss_o[1] + loss_o[2] + 10 * loss_o[3] + loss_o[4]).backward()
        lerrs = [abs(float(a) - float(b)) / max(abs(float(b)), 1e-6) for a, b in zip(loss_o, loss_r[:5])]
        grads_r = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        gerrs = {k: rel(sdo[k].grad, g) for k, g in grads_r.items() if sdo[k].grad is not None}
        print(f"[closed {tag}] oracle vs reference: outputs {max(errs.values()):.2e} losses {max(lerrs):.2e} grads {max(gerrs.values()):.2e}")
        assert max(lerrs) < 5e-5 and max(gerrs.values()) < 5e-2
        fp = {}
        for k, g in grads_r.items():
            fp[k] = [float(g.sum()), float(g.abs().sum())]
            arrays[f"{tag}.grad.{k}"] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy().copy()
        summary[tag] = dict(losses=[float(v) for v in loss_r[:5]], cost=float(cost_r), grad_fingerprint=fp)
    np.savez_compressed(os.path.join(GOLD, "model_thumos_closed.npz"), **arrays)
    with open(os.path.join(GOLD, "model_thumos_closed.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    ns.restore_cuda()


def window_case_tables(seed: int):
    """Synthetic video_infos / video_annos tables (sampled-frame units, fractional boundaries like the csv * ratio)."""
    import random
    r = random.Random(seed)
    infos, annos = {}, {}
    for v in range(6):
        name = f"video_{seed}_{v}"
        count = r.choice([120, 256, 300, 517, 1000, 1444])
        infos[name] = dict(fps=30.0, sample_fps=10.0, count=count * 3, sample_count=count)
        n = r.randint(1, 5)
        segs = []
        for _ in range(n):
            s = r.uniform(0, count - 12)
            segs.append([s, min(s + r.uniform(4, 180), count - 1.0), r.randint(1, 20)])
        annos[name] = segs
    return infos, annos


def window_cases():
    """Sliding-window index + start/end score maps: the reference's split_videos (thumos_dataset.py:69-130) vs
    opental_b200.windows.split_videos on synthetic tables."""
    import importlib
    ns = ref_loader.load_reference()
    ds = importlib.import_module("AFSD.common.thumos_dataset")
    from opental_b200 import windows as Wn
    out = []
    for seed, (clip, stride) in enumerate([(256, 30), (256, 128), (128, 30), (512, 64)]):
        infos, annos = window_case_tables(seed)
        tl_r, th_r = ds.split_videos(infos, annos, clip, stride)
        tl_o, th_o = Wn.split_videos(infos, annos, clip, stride)
        assert th_r == th_o and len(tl_r) == len(tl_o) > 0, (seed, len(tl_r), len(tl_o))
        for a, b in zip(tl_r, tl_o):
            assert a["video_name"] == b["video_name"] and a["offset"] == b["offset"] and a["annos"] == b["annos"]
            assert np.array_equal(a["start"], b["start"]) and np.array_equal(a["end"], b["end"])
        out.append(dict(seed=seed, clip_length=clip, stride=stride, th=th_r, infos=infos, annos=annos,
                        windows=[dict(video_name=w["video_name"], offset=int(w["offset"]), annos=[list(map(float, x)) for x in w["annos"]],
                                      start=np.nonzero(w["start"])[0].tolist(), end=np.nonzero(w["end"])[0].tolist()) for w in tl_r]))
        print(f"[windows] seed {seed}: {len(tl_r)} windows, th {sorted(th_r.values())}: identical")
    with open(os.path.join(GOLD, "window_cases.json"), "w") as fh:
        json.dump(out, fh)
    ns.restore_cuda()


def dataset_case_files(root: str, seed: int = 0, n_videos: int = 4):
    """Writes a miniature THUMOS14 tree in the reference's file formats (thumos_dataset.py:13-56,133-141; video2npy.py:61-74)
    and returns (info csv, annotation csv, class txt, npy dir).  Deterministic in `seed`: tests regenerate the same files."""
    import random
    r = random.Random(7000 + seed)
    npy = os.path.join(root, "npy")
    os.makedirs(npy, exist_ok=True)
    origin = [7, 9, 12, 21, 22, 23, 24, 26, 31, 33, 36, 40, 45, 51, 68, 79, 85, 92, 93, 97]
    with open(os.path.join(root, "classes.txt"), "w") as fh:
        for i, o in enumerate(origin):
            fh.write(f"{o} Class{i}\n")
    info_rows, anno_rows = [], []
    for v in range(n_videos):
        name = f"video_test_{seed:02d}{v:03d}"
        sample_count = r.choice([200, 300, 345, 410])
        count = sample_count * 3 - r.randint(0, 2)
        info_rows.append(f"{name},30.0,10.0,{count},{sample_count}")
        t = r.uniform(20, 90)
        while t + 60 < count:
            length = r.uniform(70, 330)
            end = min(t + length, count - 3.0)
            oi = r.choice(origin)
            anno_rows.append(f"{name},Class,{oi},{t / 30.0:.1f},{end / 30.0:.1f},{int(t)},{int(end)}")
            t = end + r.uniform(60, 260)
        g = torch.Generator().manual_seed(424242 + 1000 * seed + v)
        np.save(os.path.join(npy, name + ".npy"), torch.randint(0, 256, (sample_count, 112, 112, 3), generator=g, dtype=torch.uint8).numpy())
    info, anno = os.path.join(root, "info.csv"), os.path.join(root, "anno.csv")
    with open(info, "w") as fh:
        fh.write("video,fps,sample_fps,count,sample_count\n" + "\n".join(info_rows) + "\n")
    with open(anno, "w") as fh:
        fh.write("video,type,type_idx,start,end,startFrame,endFrame\n" + "\n".join(anno_rows) + "\n")
    return info, anno, os.path.join(root, "classes.txt"), npy


def dataset_cases():
    """The training set end to end: csv / txt / npy parsing, window index and `THUMOS_Dataset.__getitem__` of the reference
    (thumos_dataset.py:13-56,133-275) vs opental_b200.dataset on a miniature tree, same `random` seed on both sides.  The
    reference returns fp32 clips; ours returns the uint8 window + (crop, mirror, frame map) — `host_clip` states what the
    ingest kernel makes of them, and must equal the reference's clips bit for bit."""
    import importlib
    import random
    import tempfile
    import zlib
    ns = ref_loader.load_reference()
    ds = importlib.import_module("AFSD.common.thumos_dataset")
    from opental_b200 import dataset as D
    out = []
    for seed, training in ((0, True), (1, True), (2, False)):
        with tempfile.TemporaryDirectory() as root:
            info, anno, cls, npy = dataset_case_files(root, seed)
            vi_r = ds.get_video_info(info)
            va_r = ds.get_video_anno(vi_r, anno, cls)
            vi_o = D.get_video_info(info)
            va_o = D.get_video_anno(vi_o, anno, cls)
            assert {k: {a: float(b) for a, b in v.items()} for k, v in vi_r.items()} == \
                   {k: {a: float(b) for a, b in v.items()} for k, v in vi_o.items()}
            assert {k: [list(map(float, a)) for a in v] for k, v in va_r.items()} == \
                   {k: [list(map(float, a)) for a in v] for k, v in va_o.items()}
            assert ds.get_class_index_map(cls) == D.get_class_index_map(cls)
            data_r = ds.load_video_data(vi_r, npy)
            data_o = D.load_video_data(vi_o, npy)
            ref = ds.THUMOS_Dataset(data_r, vi_r, va_r, clip_length=256, crop_size=96, stride=30, training=training)
            ours = D.ThumosWindows(data_o, vi_o, va_o, clip_length=256, crop_size=96, stride=30, training=training)
            assert len(ref) == len(ours) > 0
            samples = []
            for idx in range(0, len(ref), max(1, len(ref) // 6)):
                random.seed(100 * seed + idx)
                clip, target, scores, ssl_clip, ssl_target, flag = ref[idx]
                s = ours.sample(idx, random.Random(100 * seed + idx))
                mine = D.host_clip(s["frames"], s["crop"], 96)
                mine_ssl = D.host_clip(s["frames"], s["crop"], 96, s["frame_map"])
                assert torch.equal(mine, clip) and torch.equal(mine_ssl, ssl_clip), (seed, idx)
                assert bool(flag) == s["flag"]
                assert np.array_equal(np.asarray(target, dtype=np.float32), s["target"])
                assert torch.equal(scores, torch.from_numpy(s["scores"]))
                assert np.array_equal(np.asarray(ssl_target, dtype=np.float32)[:, :2], s["ssl_target"])
                samples.append(dict(idx=idx, rng_seed=100 * seed + idx, crop=list(map(int, s["crop"])), flag=bool(flag),
                                    target=np.asarray(target, dtype=np.float64).tolist(),
                                    ssl_target=np.asarray(ssl_target, dtype=np.float64).tolist(),
                                    frame_map_crc=zlib.crc32(np.asarray(s["frame_map"], dtype=np.int32).tobytes()),
                                    clip_crc=zlib.crc32(clip.contiguous().numpy().tobytes()),
                                    ssl_clip_crc=zlib.crc32(ssl_clip.contiguous().numpy().tobytes()),
                                    scores_crc=zlib.crc32(scores.contiguous().numpy().tobytes())))
            out.append(dict(seed=seed, training=training, n_windows=len(ref),
                            video_infos={k: {a: float(b) for a, b in v.items()} for k, v in vi_r.items()},
                            video_annos={k: [list(map(float, a)) for a in v] for k, v in va_r.items()}, samples=samples))
            print(f"[dataset] seed {seed} (training={training}): {len(ref)} windows, {len(samples)} samples "
                  f"({sum(x['flag'] for x in samples)} augmented): clips, ssl clips, targets, scores identical")
    with open(os.path.join(GOLD, "dataset_cases.json"), "w") as fh:
        json.dump(out, fh)
    ns.restore_cuda()


CONFIG_CASE_YAML = """dataset:
  num_classes: 16
  class_info_path: ./data/open/split_{id:d}/classes.txt
  training:
    video_info_path: ./data/open/train_info.csv
    video_anno_path: ./data/open/split_{id:d}/train_anno.csv
    video_data_path: ./data/train_npy/
    clip_length: 256
    clip_stride: 30
    crop_size: 96
  testing:
    video_info_path: ./data/open/split_{id:d}/test_info.csv
    video_anno_path: ./data/open/split_{id:d}/test_anno.csv
    video_data_path: ./data/test_npy/
    crop_size: 96
    clip_length: 256
    clip_stride: 128
model:
  in_channels: 3
  freeze_bn: true
  freeze_bn_affine: true
  use_edl: true
  evidence: exp
  dropout: 0
  os_head: true
  backbone_model: ./weights/i3d.pt
training:
  batch_size: 1
  learning_rate: 1e-5
  weight_decay: 1e-3
  max_epoch: 25
  focal_loss: false
  edl_loss: true
  edl_config: {evidence: exp, loss_type: log, iou_aware: true, with_ibm: true, ibm_start: 10, momentum: 0.99, num_bins: 50}
  act_config: {margin: 1.0, weight: 0}
  checkpoint_path: ./ckpt/split_{id:d}/
  random_seed: 2020
testing:
  conf_thresh: 0.01
  top_k: 5000
  nms_thresh: 0.5
  nms_sigma: 0.5
  checkpoint_path: ./ckpt/split_{id:d}/checkpoint-latest.ckpt
  output_path: ./output/split_{id:d}
  output_json: detection_results.json
"""
CONFIG_CASE_ARGVS = [
    [],
    ["--open_set", "--split=0", "--lw=1", "--cw=10", "--ctw=1", "--ssl=0.001", "--piou=0.5"],
    ["--open_set", "--split=3", "--batch_size=8", "--learning_rate=2e-5", "--max_epoch=12", "--resume=4", "--ngpu=8"],
    ["--checkpoint_path=/tmp/ck", "--seed=7", "--nms_sigma=0.4", "--top_k=100", "--fusion", "--ood_scoring=uncertainty",
     "--output_json=out.json", "--exp_tag=x", "--weight_decay=0.01", "--actw=0.5", "--nms_thresh=0.6"],
]


def config_cases():
    """The reference's own `get_config` (AFSD/common/config.py:5-98) vs opental_b200.config.get_config on a synthetic yaml
    with the key structure of configs/thumos14_opental_final.yaml."""
    import importlib
    import tempfile
    ns = ref_loader.load_reference()
    ref_cfg = importlib.import_module("AFSD.common.config")
    from opental_b200 import config as C
    out = []
    with tempfile.TemporaryDirectory() as root:
        path = os.path.join(root, "cfg.yaml")
        with open(path, "w") as fh:
            fh.write(CONFIG_CASE_YAML)
        for argv in CONFIG_CASE_ARGVS:
            saved = sys.argv
            sys.argv = ["ref", path, *argv]
            try:
                want = ref_cfg.get_config()
            finally:
                sys.argv = saved
            got = C.get_config([path, *argv])
            assert got == want, (argv, got, want)
            out.append(dict(argv=argv, config=want))
            print(f"[config] argv {argv}: identical")
    with open(os.path.join(GOLD, "config_cases.json"), "w") as fh:
        json.dump(dict(yaml=CONFIG_CASE_YAML, cases=out), fh)
    ns.restore_cuda()


def anet_dataset_case_files(root: str, seed: int = 0, n_videos: int = 5, clip_length: int = 768):
    """A miniature ActivityNet tree in the reference's formats (anet_dataset.py:32-105, 226-229): video_info json + npy."""
    import random
    r = random.Random(9000 + seed)
    npy = os.path.join(root, "npy")
    os.makedirs(npy, exist_ok=True)
    info = {}
    for v in range(n_videos):
        name = f"v_{seed:02d}{v:03d}"
        frame_num = clip_length if v != 1 else 700                       # one short video: the padding path
        annos, t = [], r.uniform(10, 120)
        while t + 80 < frame_num:
            length = r.uniform(60, 330)
            end = min(t + length, frame_num - 3.0)
            annos.append(dict(start_frame=round(t, 2), end_frame=round(end, 2), label_id=r.randint(1, 150)))
            t = end + r.uniform(60, 260)
        if v == 2:
            annos.append(dict(start_frame=50.0, end_frame=50.0, label_id=3))        # dropped: end <= start
        info[name] = dict(subset="training" if v != 3 else "validation", frame_num=frame_num, annotations=annos)
        if v != 4:                                                                  # video 4 has no npy file: skipped
            g = torch.Generator().manual_seed(515151 + 1000 * seed + v)
            np.save(os.path.join(npy, name + ".npy"), torch.randint(0, 256, (frame_num, 112, 112, 3), generator=g, dtype=torch.uint8).numpy())
    path = os.path.join(root, "video_info.json")
    with open(path, "w") as fh:
        json.dump(info, fh)
    return path, npy


def anet_dataset_cases():
    """`ANET_Dataset` of the reference (anet_dataset.py:127-257) vs opental_b200.anet_dataset on a miniature tree, same `random`
    seed.  numpy >= 1.24 has no `np.float` (used at :232): aliased for the duration of the run (test infrastructure only)."""
    import importlib
    import random
    import tempfile
    import zlib
    ns = ref_loader.load_reference()
    ds = importlib.import_module("AFSD.common.anet_dataset")
    from opental_b200 import anet_dataset as AD
    from opental_b200 import dataset as D
    if not hasattr(np, "float"):
        np.float = float
    out = []
    for seed, training in ((0, True), (1, True), (2, False)):
        with tempfile.TemporaryDirectory() as root:
            info_path, npy = anet_dataset_case_files(root, seed)
            subset = "training" if training else "validation"
            ref = ds.ANET_Dataset(info_path, npy, 768, 96, 768, training=training)
            ours = AD.AnetWindows(info_path, npy, 768, 96, training=training)
            assert ds.get_video_info(info_path, subset) == AD.get_video_info(info_path, subset)
            assert len(ref) == len(ours) > 0 and ref.th == ours.th
            samples = []
            for idx in range(len(ref)):
                a, b = ref.training_list[idx], ours.training_list[idx]
                assert a["video_name"] == b["video_name"] and a["annos"] == b["annos"] and a["frame_num"] == b["frame_num"]
                for k in ("start", "end", "action"):
                    assert np.array_equal(a[k], b[k]), (seed, idx, k)
                random.seed(100 * seed + idx)
                clip, target, scores, ssl_clip, ssl_target, flag = ref[idx]
                s = ours.sample(idx, random.Random(100 * seed + idx))
                mine = D.host_clip(s["frames"], s["crop"], 96)
                mine_ssl = D.host_clip(s["frames"], s["crop"], 96, s["frame_map"])
                n_real = a["frame_num"]                                             # beyond: the 127.5-vs-128 padding deviation
                assert torch.equal(mine[:, :n_real], clip[:, :n_real]), (seed, idx)
                if n_real == 768:
                    assert torch.equal(mine_ssl, ssl_clip), (seed, idx)
                else:
                    assert float((mine[:, n_real:] - clip[:, n_real:]).abs().max()) <= 1.0 / 255 + 1e-6
                assert bool(flag) == s["flag"]
                assert np.array_equal(np.asarray(target, dtype=np.float32), s["target"])
                assert torch.equal(scores, torch.from_numpy(s["scores"]))
                assert np.array_equal(np.asarray(ssl_target, dtype=np.float32)[:, :2], s["ssl_target"])
                samples.append(dict(idx=idx, rng_seed=100 * seed + idx, crop=list(map(int, s["crop"])), flag=bool(flag), frame_num=int(n_real),
                                    target=np.asarray(target, dtype=np.float64).tolist(),
                                    ssl_target=np.asarray(ssl_target, dtype=np.float64).tolist(),
                                    frame_map_crc=zlib.crc32(np.asarray(s["frame_map"], dtype=np.int32).tobytes()),
                                    clip_crc=zlib.crc32(clip[:, :n_real].contiguous().numpy().tobytes()),
                                    ssl_clip_crc=zlib.crc32(ssl_clip.contiguous().numpy().tobytes()) if n_real == 768 else None,
                                    scores_crc=zlib.crc32(scores.contiguous().numpy().tobytes())))
            out.append(dict(seed=seed, training=training, n_windows=len(ref), th=ref.th, samples=samples))
            print(f"[anet dataset] seed {seed} (training={training}): {len(ref)} windows ({sum(x['flag'] for x in samples)} augmented): identical")
    with open(os.path.join(GOLD, "anet_dataset_cases.json"), "w") as fh:
        json.dump(out, fh)
    ns.restore_cuda()


def augment_cases_anet():
    """The ActivityNet variant of the cut-paste augmentation (anet_dataset.py:171-221: `>=` length test, RuntimeError of a
    mis-sized slice assignment = not augmented) — the reference's own method on a frame-index clip, as augment_cases()."""
    import importlib
    import random
    import types
    ns = ref_loader.load_reference()
    ds = importlib.import_module("AFSD.common.anet_dataset")
    from opental_b200 import augment as A
    dummy = types.SimpleNamespace(clip_length=256)
    dummy.get_bg = types.MethodType(ds.ANET_Dataset.get_bg, dummy)
    dummy.augment_ = types.MethodType(ds.ANET_Dataset.augment_, dummy)
    clip = torch.arange(256, dtype=torch.float32).view(1, 256, 1, 1).expand(3, 256, 2, 2).contiguous()
    cases, n_ok, n_crash = [], 0, 0
    for seed in range(160):
        annos, th = augment_case_inputs(seed)
        if seed % 3 == 0:                                   # actions of exactly 2*th frames / pastes near the clip end
            annos = [[a[0], a[0] + 2 * th, a[2]] for a in annos[:1]] + annos[1:]
        random.seed(seed)
        try:
            new_input, new_annos, flag = ds.ANET_Dataset.augment(dummy, clip, [list(a) for a in annos], th, 1)
        except IndexError:
            # the reference itself crashes (random.choice of an empty range when an action is exactly 2*th long, :179-180);
            # ours must raise the same error
            random.seed(seed)
            try:
                A.cut_paste([list(a) for a in annos], th, 256, 1, variant="anet")
                raise AssertionError(f"seed {seed}: the reference raised IndexError, ours did not")
            except IndexError:
                n_crash += 1
            continue
        ref_map = new_input[0, :, 0, 0].long().tolist()
        random.seed(seed)
        fmap, got_annos, got_flag = A.cut_paste([list(a) for a in annos], th, 256, 1, variant="anet")
        assert got_flag == flag and fmap.tolist() == ref_map, seed
        assert [list(map(float, a)) for a in got_annos] == [list(map(float, a)) for a in new_annos], seed
        n_ok += bool(flag)
        cases.append(dict(seed=seed, annos=annos, th=th, flag=bool(flag), frame_map=ref_map,
                          new_annos=[list(map(float, a)) for a in new_annos]))
    print(f"[augment anet] {len(cases)} cases ({n_ok} augmented, {n_crash} seeds where reference and ours both raise IndexError)")
    with open(os.path.join(GOLD, "augment_cases_anet.json"), "w") as fh:
        json.dump(cases, fh)
    ns.restore_cuda()


def anet_ssl_cases():
    """The SSL / triplet second pass of the ActivityNet flavour (anet/BDNet.py:453-474, anet/train.py:159-166, 222-226): the
    reference's own forward(ssl=True) + triplet loss + backward on the keyed synthetic weights and a 768-frame clip."""
    import torch.nn as nn
    ns = ref_loader.load_reference(config="configs/anet_opental.yaml", extra_args=("--open_set", "--split=0"), flavour="anet")
    cfg = O.anet_config()
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=math.log(8.0))
    net = ns.BDNet(in_channels=3, training=False, frame_num=768, use_edl=True)
    net.load_state_dict(sd)
    net.train()
    x = O.synthetic_clip(1, frames=768).unsqueeze(0)
    proposals = [torch.tensor([[120.0, 270.0], [366.0, 516.0], [273.0, 363.0]])]      # anchor / positive / negative, frames
    net.zero_grad()
    a_r, p_r, n_r = net(x, proposals=proposals, ssl=True)
    trip_r = torch.stack([nn.TripletMarginLoss()(a_r[i], p_r[i], n_r[i]) * w for i, w in enumerate((1, 0.1, 0.1))]).sum(0)
    trip_r.backward()
    grads_r = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    arrays = {}
    for name, lst in (("anchor", a_r), ("positive", p_r), ("negative", n_r)):
        for i, t in enumerate(lst):
            arrays[f"ssl.{name}.{i}"] = t.detach().numpy()
    fp = {}
    for k, g in grads_r.items():
        fp[k] = [float(g.sum()), float(g.abs().sum())]
        arrays[f"ssl.grad.{k}"] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy().copy()
    print(f"[anet ssl] reference: triplet {float(trip_r):.5f}, {len(fp)} parameter gradients, feature shapes {[tuple(t.shape) for t in a_r]}")
    np.savez_compressed(os.path.join(GOLD, "model_anet_ssl.npz"), **arrays)
    with open(os.path.join(GOLD, "model_anet_ssl.json"), "w") as fh:
        json.dump(dict(triplet=float(trip_r), proposals=proposals[0].tolist(), grad_fingerprint=fp), fh, indent=1)
    ns.restore_cuda()


def batch8_cases():
    """The BENCH configuration (BASELINE configs[1]/[2]: batch 8 per GPU) as a golden: reference BDNet forward +
    MultiSegmentLoss (epoch 1 and 11) + backward on synthetic clips 0..7 — the whole-model path at B = 8 exercises the
    `[P,B]`-vs-`[B,P]` IoU-calibration pairing (App. D) that B = 1 cannot.  `--batch8` (about a minute, ~25 GB)."""
    ns = ref_loader.load_reference()
    cfg = O.OracleConfig()
    B = 8
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=math.log(32.0))
    net = ns.BDNet(in_channels=3, training=False, use_edl=True)
    net.load_state_dict(sd)
    net.train()
    x = torch.stack([O.synthetic_clip(i) for i in range(B)])
    targets = [O.synthetic_targets(i, num_classes=cfg.num_classes) for i in range(B)]
    targets[3] = targets[3][:1].clone()                                   # ragged: 1, 2 and 3 segments per clip
    targets[5] = torch.cat([targets[5], torch.tensor([[0.42, 0.47, 3.0]])])
    summary, arrays = dict(n_segments=[int(t.shape[0]) for t in targets]), {}
    net.zero_grad()
    out_r = net(x)
    with torch.no_grad():
        out_o = O.bdnet_forward(x, sd, cfg, compat=True)
    errs = {k: rel(out_o[k].detach(), out_r[k].detach()) for k in out_r if out_r[k] is not None}
    # Everything downstream of the proposal windows is a DISCONTINUOUS function of `loc` (BDNet.py:355-384 rounds the window
    # ends to frames): with 8 x 126 x 4 window ends, a 3e-6 difference in `loc` moves one of them across an integer and the
    # pooled maximum of that prior changes by ~5e-4 of the output range.  So: the coarse outputs at TOL, the refined ones
    # row-wise — all but a handful of (clip, prior) rows at TOL.
    refined = ("prop_loc", "prop_conf", "center", "prop_act", "prop_unct")
    assert max(v for k, v in errs.items() if k not in refined) < TOL, errs
    flipped = set()
    for k in refined:
        d = (out_o[k].detach() - out_r[k].detach()).abs().reshape(B, 126, -1).amax(-1) / out_r[k].detach().abs().max()
        flipped |= {tuple(i) for i in (d > TOL).nonzero().tolist()}
        assert errs[k] < 2e-3, errs
    assert len(flipped) <= 4, flipped
    summary["oracle_vs_ref_flipped_rows"] = sorted(flipped)
    for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "unct", "prop_unct"):
        arrays[f"b8.{k}"] = out_r[k].detach().numpy()
    for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
        arrays[f"b8.{k}.sample"] = out_r[k].detach().numpy()[:, ::8, ::8].copy()
    for epoch in (1, 11):
        crit = ns.MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=ns.config["training"]["edl_config"],
                                   os_head=True, act_config=ns.config["training"]["act_config"])
        crit.cls_loss.epoch = epoch
        loss_r = crit(out_r, [t.clone() for t in targets])
        state = O.LossState(epoch=epoch)
        loss_o = O.multisegment_loss({k: (v.detach() if torch.is_tensor(v) else v) for k, v in out_r.items()}, targets, state, cfg)
        lerrs = [abs(float(a) - float(b)) / max(abs(float(b)), 1e-6) for a, b in zip(loss_o, loss_r)]
        assert max(lerrs) < 5e-5, lerrs
        assert torch.allclose(crit.cls_loss.weight_accum, state.weight_accum, atol=1e-6)
        cost_r = loss_r[0] + 10 * loss_r[1] + loss_r[2] + 10 * loss_r[3] + loss_r[4] + loss_r[5] + loss_r[6]
        summary[f"e{epoch}"] = dict(losses=[float(v) for v in loss_r], cost=float(cost_r))
        arrays[f"b8.e{epoch}.weight_accum"] = crit.cls_loss.weight_accum.numpy().copy()
    cost_r.backward()                                                     # epoch-11 cost
    fp = {}
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        g = p.grad
        fp[k] = [float(g.sum()), float(g.abs().sum())]
        arrays[f"b8.e11.grad.{k}"] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy().copy()
    summary["e11"]["grad_fingerprint"] = fp
    print(f"[batch8] oracle vs reference outputs {max(errs.values()):.2e}; losses e1 {summary['e1']['losses']} e11 {summary['e11']['losses']}")
    np.savez_compressed(os.path.join(GOLD, "model_thumos_b8.npz"), **arrays)
    with open(os.path.join(GOLD, "model_thumos_b8.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    ns.restore_cuda()


TRAJ_STEPS, TRAJ_LR, TRAJ_WD = 5, 1e-5, 1e-3


def trajectory_cases():
    """A 5-step training trajectory of the reference's own modules under `torch.optim.Adam(net.parameters(), lr,
    weight_decay)` (thumos14/train.py:226-252, 321-323): batch of 2 clips, epoch 11 (IBM on: the 50-bin EMA evolves from step
    to step), cost incl. the boundary BCE terms of forward_one_epoch (train.py:186-200, restated here because importing
    train.py needs tensorboardX and creates directories).  lr = 1e-5, weight decay 1e-3 (the config's values): the cost falls 77.2 -> ~70 in five steps.  Stores cost / losses per step and the per-tensor |delta w| sums.  `--trajectory`."""
    ns = ref_loader.load_reference()
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=math.log(32.0))
    net = ns.BDNet(in_channels=3, training=False, use_edl=True)
    net.load_state_dict(sd)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=TRAJ_LR, weight_decay=TRAJ_WD)
    crit = ns.MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=ns.config["training"]["edl_config"],
                               os_head=True, act_config=ns.config["training"]["act_config"])
    crit.cls_loss.epoch = 11
    x = torch.stack([O.synthetic_clip(i) for i in range(2)])
    targets = [O.synthetic_targets(i, num_classes=cfg.num_classes) for i in range(2)]
    scores = torch.stack([O.synthetic_scores(t) for t in targets])
    # the oracle runs the same trajectory (functional weights + its own Adam)
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "num_batches" not in k else v)
           for k, v in sd.items()}
    train_keys = [k for k, p in net.named_parameters() if p.requires_grad]
    opt_o = torch.optim.Adam([sdo[k] for k in train_keys], lr=TRAJ_LR, weight_decay=TRAJ_WD)
    state = O.LossState(epoch=11)
    steps = []
    for s in range(TRAJ_STEPS):
        out = net(x)
        l, c, pl, pc, ct, act, pact = crit(out, [t.clone() for t in targets])
        ls, le = O.boundary_bce(out["start"], out["end"], scores)
        sc4 = torch.nn.functional.interpolate(scores, scale_factor=1.0 / 4)
        a, b = O.boundary_bce(out["start_loc_prop"], out["end_loc_prop"], sc4)
        c2, d = O.boundary_bce(out["start_conf_prop"], out["end_conf_prop"], sc4)
        ls = ls + 0.1 * (a + c2); le = le + 0.1 * (b + d)
        cost = l + 10 * c + pl + 10 * pc + ct + ls + le + act + pact
        opt.zero_grad()
        cost.backward()
        opt.step()
        out_o = O.bdnet_forward(x, sdo, cfg, compat=True)
        cost_o, parts = O.training_cost(out_o, targets, scores, state, cfg)
        opt_o.zero_grad()
        cost_o.backward()
        opt_o.step()
        e = abs(float(cost_o) - float(cost)) / abs(float(cost))
        print(f"[trajectory] step {s}: reference cost {float(cost):.6f}  oracle {float(cost_o):.6f}  rel {e:.2e}")
        # the trajectory is chaotic beyond a few steps (matching flips a prior when an IoU crosses 0.5, Adam's first steps are
        # sign-like): two fp32 CPU implementations agree to < 1e-4 for three steps and drift to ~1e-3 afterwards.  The drift
        # is recorded; tests allow max(1e-3, 3 x drift) at each step.
        assert s > 2 or e < 2e-4, (s, float(cost), float(cost_o))
        steps.append(dict(cost=float(cost), oracle_rel=e, losses=[float(v) for v in (l, c, pl, pc, ct, act, pact)], loss_start=float(ls),
                          loss_end=float(le)))
    delta = {k: float((p.detach() - sd[k]).abs().sum()) for k, p in net.named_parameters() if p.requires_grad}
    arrays = {"weight_accum": crit.cls_loss.weight_accum.numpy().copy()}
    with open(os.path.join(GOLD, "trajectory_thumos.json"), "w") as fh:
        json.dump(dict(steps=steps, lr=TRAJ_LR, weight_decay=TRAJ_WD, delta_abs_sum=delta,
                       delta_total=float(sum(delta.values()))), fh, indent=1)
    np.savez_compressed(os.path.join(GOLD, "trajectory_thumos.npz"), **arrays)
    ns.restore_cuda()


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if "--batch8" in sys.argv:
        batch8_cases()
    elif "--trajectory" in sys.argv:
        trajectory_cases()
    elif "--anet" in sys.argv:
        anet_cases()
    elif "--ssl" in sys.argv:
        ssl_cases()
    elif "--infer" in sys.argv:
        infer_cases()
    elif "--edl" in sys.argv:
        edl_cases()
    elif "--closed" in sys.argv:
        closed_cases()
    elif "--windows" in sys.argv:
        sys.path.insert(0, ROOT)
        window_cases()
    elif "--anet-ssl" in sys.argv:
        anet_ssl_cases()
    elif "--augment-anet" in sys.argv:
        sys.path.insert(0, ROOT)
        augment_cases_anet()
    elif "--anet-dataset" in sys.argv:
        sys.path.insert(0, ROOT)
        anet_dataset_cases()
    elif "--config" in sys.argv:
        sys.path.insert(0, ROOT)
        config_cases()
    elif "--dataset" in sys.argv:
        sys.path.insert(0, ROOT)
        dataset_cases()
    elif "--augment" in sys.argv:
        sys.path.insert(0, ROOT)
        augment_cases()
    else:
        bmp_cases()
        model_cases()
    print("golden fixtures written to", GOLD)
