"""CPU: the host-side mirror of the reference's module API — constructor arguments from the reference's own yaml files,
state_dict keys / shapes identical to the oracle's specification of the reference, for both the
THUMOS14 and the ActivityNet flavour.  (No kernel runs: construction and parameter bookkeeping are plain torch.)"""
import os

import pytest
import torch

import opental_oracle as O

REF = "/root/reference"


def _spec(cfg):
    return {k: tuple(s) for k, s, _ in O.model_spec(cfg)}


@pytest.mark.parametrize("variant", ["thumos", "anet"])
def test_state_dict_matches_reference_spec(variant):
    from opental_b200.bdnet import BDNet
    if variant == "thumos":
        cfg, net = O.OracleConfig(), BDNet(training=False, use_edl=True, num_classes=16, os_head=True)
    else:
        cfg, net = O.anet_config(), BDNet(training=False, use_edl=True, num_classes=151, os_head=True, frame_num=768, variant="anet")
    sd = net.state_dict()
    spec = _spec(cfg)
    assert set(sd) == set(spec)
    assert all(tuple(sd[k].shape) == spec[k] for k in spec)
    # loading keyed synthetic weights and reading them back is lossless (parameters are views into flat buffers)
    src = O.synthetic_state_dict(cfg)
    net.load_state_dict(src)
    back = net.state_dict()
    assert all(torch.equal(back[k], src[k]) for k in src)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
@pytest.mark.parametrize("yaml_name,variant,classes,priors", [("thumos14_opental_final.yaml", "thumos", 15, 126),
                                                               ("anet_opental.yaml", "anet", 150, 189),
                                                               ("thumos14.yaml", "thumos", 21, 126)])
def test_from_config_reads_the_reference_yaml(yaml_name, variant, classes, priors):
    import yaml
    from opental_b200.bdnet import BDNet
    with open(os.path.join(REF, "configs", yaml_name)) as fh:
        cfg = yaml.safe_load(fh)
    # the reference's own constructor call (AFSD/thumos14/train.py:314) plus the flavour switch
    net = BDNet.from_config(cfg, in_channels=cfg["model"]["in_channels"], backbone_model=None, training=False,
                            use_edl=cfg["model"].get("use_edl", False), variant=variant)
    assert net.num_classes == classes and net.coarse_pyramid_detection.num_priors == priors
    assert net.coarse_pyramid_detection.frame_num == cfg["dataset"]["training"]["clip_length"]


def test_unsupported_variants_raise():
    from opental_b200.bdnet import BDNet
    with pytest.raises(NotImplementedError):
        BDNet(training=False, use_rpl=True)
    with pytest.raises(NotImplementedError):
        BDNet(training=False, dropout=0.1)
    from opental_b200.multisegment_loss import MultiSegmentLoss
    with pytest.raises(NotImplementedError):
        MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="rpl")


def test_optimizer_state_interoperates_with_torch_adam():
    """The flat Adam moment buffers <-> `torch.optim.Adam(net.parameters()).state_dict()` (the 'optimizer' entry of the
    reference's training-state file, train.py:115), through the parameters' views into the flat buffers."""
    from opental_b200 import checkpoint as ck
    from opental_b200.bdnet import BDNet
    from opental_b200.engine import FlatParams
    net = BDNet(training=False, use_edl=True, num_classes=16, os_head=True)
    net.train()
    dev = torch.device("cpu")
    bb, hc = net.backbone, net.coarse_pyramid_detection.conv_store
    w, g = bb.flat_parameters(dev)
    hc.ensure(dev)
    skip = {id(p) for p in bb.parameters()} | {id(r.weight) for r in hc.recs}
    head = FlatParams([p for p in net.parameters() if p.requires_grad and id(p) not in skip])
    groups = [(w, g), (hc.flat_w, hc.flat_g), (head.w, head.g)]
    gen = torch.Generator().manual_seed(0)
    state = [dict(m=torch.randn(a.shape, generator=gen), v=torch.rand(a.shape, generator=gen)) for a, _ in groups]
    params = list(net.parameters())
    sd = ck.adam_state_dict(params, groups, state, step=7, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3)
    trainable = [i for i, p in enumerate(params) if p.requires_grad]
    assert sorted(sd["state"]) == trainable and sum(params[i].numel() for i in trainable) == 44721259
    opt = torch.optim.Adam(net.parameters(), lr=1.0)
    opt.load_state_dict(sd)                                  # torch accepts it ...
    back = opt.state_dict()
    assert back["param_groups"][0]["lr"] == 1e-5 and back["param_groups"][0]["weight_decay"] == 1e-3
    state2 = [dict(m=torch.zeros_like(a), v=torch.zeros_like(a)) for a, _ in groups]
    assert ck.load_adam_state_dict(back, params, groups, state2) == 7      # ... and its own dict loads back losslessly
    for i in trainable:
        for key in ("m", "v"):
            a = ck.moment_view(params[i], groups, [s[key] for s in state])
            assert torch.equal(a, ck.moment_view(params[i], groups, [s[key] for s in state2]))
            assert torch.equal(a, sd["state"][i]["exp_avg" if key == "m" else "exp_avg_sq"])
    # torch 1.9 (the reference's pin) stores `step` as a python int
    for ent in back["state"].values():
        ent["step"] = 7
    assert ck.load_adam_state_dict(back, params, groups, state2) == 7
    with pytest.raises(ValueError):
        bad = dict(state={}, param_groups=[dict(params=list(range(3)))])
        ck.load_adam_state_dict(bad, params, groups, state2)


def test_two_rate_optimizer_state_matches_the_activitynet_layout():
    """anet/train.py:304-311 builds Adam from two groups (backbone at 0.1 x the rate, then coarse_pyramid_detection): parameters
    are numbered backbone-first, unlike net.parameters().  The state dict in that layout loads into such a torch optimizer and
    comes back; a single-rate file is refused with a message instead of being mis-assigned."""
    from opental_b200 import checkpoint as ck
    from opental_b200.bdnet import BDNet
    from opental_b200.engine import FlatParams
    net = BDNet(training=False, use_edl=True, num_classes=16, os_head=True)
    net.train()
    dev = torch.device("cpu")
    bb, hc = net.backbone, net.coarse_pyramid_detection.conv_store
    w, g = bb.flat_parameters(dev)
    hc.ensure(dev)
    skip = {id(p) for p in bb.parameters()} | {id(r.weight) for r in hc.recs}
    head = FlatParams([p for p in net.parameters() if p.requires_grad and id(p) not in skip])
    groups = [(w, g), (hc.flat_w, hc.flat_g), (head.w, head.g)]
    gen = torch.Generator().manual_seed(1)
    state = [dict(m=torch.randn(a.shape, generator=gen), v=torch.rand(a.shape, generator=gen)) for a, _ in groups]
    params = list(net.parameters())
    pg = [(list(net.backbone.parameters()), 1e-6), (list(net.coarse_pyramid_detection.parameters()), 1e-5)]
    sd = ck.adam_state_dict(params, groups, state, step=3, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, param_groups=pg)
    assert [len(x["params"]) for x in sd["param_groups"]] == [len(pg[0][0]), len(pg[1][0])]
    assert [x["lr"] for x in sd["param_groups"]] == [1e-6, 1e-5]
    opt = torch.optim.Adam([{"params": net.backbone.parameters(), "lr": 1.0}, {"params": net.coarse_pyramid_detection.parameters(), "lr": 1.0}])
    opt.load_state_dict(sd)
    back = opt.state_dict()
    # index 0 is the backbone's first parameter in this layout
    first_bb = pg[0][0][0]
    if first_bb.requires_grad:
        assert torch.equal(back["state"][0]["exp_avg"], ck.moment_view(first_bb, groups, [s["m"] for s in state]))
    state2 = [dict(m=torch.zeros_like(a), v=torch.zeros_like(a)) for a, _ in groups]
    assert ck.load_adam_state_dict(back, params, groups, state2, param_groups=pg) == 3
    for key in ("m", "v"):
        for p in params:
            if p.requires_grad:
                assert torch.equal(ck.moment_view(p, groups, [s[key] for s in state]), ck.moment_view(p, groups, [s[key] for s in state2]))
    one = ck.adam_state_dict(params, groups, state, step=3, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    with pytest.raises(ValueError, match="not interchangeable"):
        ck.load_adam_state_dict(one, params, groups, state2, param_groups=pg)


def test_proposal_list_has_the_reference_format():
    """inference.to_proposal_list: the dict layout of get_video_detections (test.py:182-200)."""
    from opental_b200.inference import results_json, to_proposal_list
    result = {2: torch.tensor([[1.0, 2.5, 0.3, 0.1, 0.9], [4.0, 6.0, 0.8, 0.2, 0.7]]), 0: torch.tensor([[0.5, 1.5, 0.6, 0.4, 0.95]])}
    names = {i: f"class{i}" for i in range(1, 16)}
    props = to_proposal_list(result, names)
    assert [p["label"] for p in props] == ["class1", "class3", "class3"]              # model class c -> name of c + 1 (os_head)
    assert [round(p["score"], 6) for p in props] == [0.6, 0.8, 0.3]                   # descending inside a class
    assert props[1]["segment"] == [4.0, 6.0] and set(props[0]) == {"label", "score", "segment", "uncertainty", "actionness"}
    closed = to_proposal_list({3: result[0]}, names, os_head=False, use_edl=False)
    assert closed[0]["label"] == "class3" and closed[0]["uncertainty"] == 0.0 and closed[0]["actionness"] == 0.0
    doc = results_json({"video_test_0000004": props})
    assert doc["version"] == "THUMOS14" and list(doc["results"]) == ["video_test_0000004"] and doc["external_data"] == {}
