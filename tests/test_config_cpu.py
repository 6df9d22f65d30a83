"""CPU: the reference's command line + yaml handling (AFSD/common/config.py:5-98) through opental_b200.config, against the
dicts its own `get_config` produced for the same argv (oracle/make_golden.py --config)."""
import json
import os

import pytest

from opental_b200 import config as C


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "config_cases.json")) as fh:
        return json.load(fh)


def test_get_config_matches_reference(golden, tmp_path):
    path = tmp_path / "cfg.yaml"
    path.write_text(golden["yaml"])
    assert len(golden["cases"]) == 4
    for case in golden["cases"]:
        assert C.get_config([str(path), *case["argv"]]) == case["config"], case["argv"]


def test_open_set_split_goes_into_the_paths(golden, tmp_path):
    path = tmp_path / "cfg.yaml"
    path.write_text(golden["yaml"])
    cfg = C.get_config([str(path), "--open_set", "--split=2"])
    assert cfg["dataset"]["class_info_path"].endswith("split_2/classes.txt")
    assert cfg["dataset"]["training"]["video_info_path"].endswith("open/train_info.csv")        # no 'split_' in it: untouched
    assert cfg["dataset"]["testing"]["video_info_path"].endswith("split_2/test_info.csv")
    assert cfg["training"]["checkpoint_path"].endswith("split_2/") and cfg["testing"]["output_path"].endswith("split_2")
    closed = C.get_config([str(path)])
    assert "{id:d}" in closed["dataset"]["class_info_path"] and closed["open_set"] is False


def test_loss_arguments_follow_train_script(golden, tmp_path):
    path = tmp_path / "cfg.yaml"
    path.write_text(golden["yaml"])
    cfg = C.get_config([str(path), "--open_set", "--piou=0.5"])
    kw = C.loss_arguments(cfg)
    assert kw["num_classes"] == 15 and kw["os_head"] and kw["cls_loss_type"] == "edl" and kw["overlap_thresh"] == 0.5
    assert kw["edl_config"]["ibm_start"] == 10 and kw["act_config"] == {"margin": 1.0, "weight": 0}
    assert C.loss_arguments(cfg, script_compat=True)["cls_loss_type"] == "focal"              # train.py:31 as written (D2)
    cfg["training"]["rpl_loss"] = True
    assert C.loss_arguments(cfg)["cls_loss_type"] == "rpl"


def test_bdnet_from_config_reads_the_same_dict(golden, tmp_path):
    from opental_b200.bdnet import BDNet
    path = tmp_path / "cfg.yaml"
    path.write_text(golden["yaml"])
    cfg = C.get_config([str(path), "--open_set"])
    net = BDNet.from_config(cfg, backbone_model=None, training=False, use_edl=True)
    assert net.os_head and net.num_classes == 15 and net.use_edl
