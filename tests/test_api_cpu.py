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
