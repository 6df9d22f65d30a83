"""CPU: tools/train_thumos.py end to end against the C-ABI emulation — the reference's command line and yaml, the dataset files,
the pretrained-backbone file, BDNet.from_config, the loss arguments, Trainer, loader threads, train_loop.fit and the epoch
log — on a miniature dataset, two steps of one epoch (`--steps_per_epoch`, `--device cpu` exist for this harness: without
the emulation the product raises on a CPU device)."""
import importlib.util
import json
import os

import pytest
import torch

import abi_emu
import make_golden
import opental_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_train_thumos_script_smoke(monkeypatch, tmp_path, capsys):
    abi_emu.install(monkeypatch)
    info, anno, cls, npy = make_golden.dataset_case_files(str(tmp_path / "data"), seed=1, n_videos=2)
    # the Kinetics I3D file the reference loads with strict=False (BDNet.py:35-37): here the synthetic backbone weights
    i3d = {k[len("backbone._model."):]: v for k, v in O.synthetic_state_dict(O.OracleConfig()).items() if k.startswith("backbone._model.")}
    torch.save(i3d, tmp_path / "i3d.pt")
    yaml_text = make_golden.CONFIG_CASE_YAML
    for old, new in (("./data/open/train_info.csv", info), ("./data/open/split_{id:d}/train_anno.csv", anno),
                     ("./data/open/split_{id:d}/classes.txt", cls), ("./data/train_npy/", npy),
                     ("./weights/i3d.pt", str(tmp_path / "i3d.pt")), ("./ckpt/split_{id:d}/", str(tmp_path / "ckpt_{id:d}") + "/")):
        assert old in yaml_text, old
        yaml_text = yaml_text.replace(old, new)
    yaml_text = yaml_text.replace("num_classes: 16", "num_classes: 21")        # the miniature tree has 20 classes
    cfg_path = tmp_path / "cfg.yaml"
    cfg_path.write_text(yaml_text)
    log = tmp_path / "log.json"
    tool = load_tool("train_thumos")
    rc = tool.main([str(cfg_path), "--open_set", "--split=0", "--lw=1", "--cw=10", "--ctw=1", "--ssl=0.001", "--piou=0.5",
                    "--max_epoch=1", "--batch_size=1", "--steps_per_epoch=2", "--device=cpu", "--no_graph", "--loader_threads=2",
                    f"--log_json={log}"])
    assert rc == 0
    text = capsys.readouterr().out
    assert "windows of 2 videos" in text and "loss edl" in text and "Epoch-1 Train Loss: Total -" in text
    hist = json.loads(log.read_text())
    assert len(hist) == 1 and hist[0]["steps"] == 2 and all(k in hist[0] for k in ("cost", "loc", "conf", "start", "end", "grad_norm"))
    assert os.path.isdir(tmp_path / "ckpt_0" / "training")


ANET_YAML = """dataset:
  num_classes: 151
  class_info_path: {root}/action_known_{{id:d}}.txt
  training:
    video_mp4_path: {npy}
    video_info_path: {info}
    video_anno_path: None
    video_data_path: None
    clip_length: 768
    clip_stride: 768
    crop_size: 96
  testing:
    video_mp4_path: {npy}
    video_info_path: {info}
    video_anno_path: None
    video_data_path: None
    crop_size: 96
    clip_length: 768
    clip_stride: 768
model:
  in_channels: 3
  freeze_bn: true
  freeze_bn_affine: true
  use_edl: true
  evidence: exp
  os_head: true
  backbone_model: {i3d}
training:
  batch_size: 1
  learning_rate: 1e-4
  weight_decay: 1e-4
  max_epoch: 25
  focal_loss: false
  edl_loss: true
  edl_config: {{evidence: exp, loss_type: log, iou_aware: true, with_ibm: true, ibm_start: 10, momentum: 0.99, num_bins: 50}}
  checkpoint_path: {root}/ckpt_{{id:d}}/
  random_seed: 2020
testing:
  conf_thresh: 0.01
  top_k: 5000
  nms_thresh: 0.5
  nms_sigma: 0.85
  checkpoint_path: {root}/ckpt_{{id:d}}/checkpoint-latest.ckpt
  output_path: {root}/out_{{id:d}}
  output_json: detection_results.json
"""


def test_train_anet_script_smoke(monkeypatch, tmp_path, capsys):
    abi_emu.install(monkeypatch)
    info, npy = make_golden.anet_dataset_case_files(str(tmp_path / "data"), seed=0, n_videos=3)
    i3d = {k[len("backbone._model."):]: v for k, v in O.synthetic_state_dict(O.OracleConfig()).items() if k.startswith("backbone._model.")}
    torch.save(i3d, tmp_path / "i3d.pt")
    cfg_path = tmp_path / "anet.yaml"
    cfg_path.write_text(ANET_YAML.format(root=str(tmp_path), npy=npy, info=info, i3d=str(tmp_path / "i3d.pt")))
    log = tmp_path / "log.json"
    rc = load_tool("train_anet").main([str(cfg_path), "--open_set", "--split=0", "--lw=1", "--cw=1", "--piou=0.6", "--max_epoch=1",
                                       "--batch_size=1", "--steps_per_epoch=1", "--device=cpu", "--no_graph", "--loader_threads=1",
                                       f"--log_json={log}"])
    assert rc == 0
    hist = json.loads(log.read_text())
    assert len(hist) == 1 and hist[0]["steps"] == 1 and hist[0]["cost"] == hist[0]["cost"]
    assert "videos;" in capsys.readouterr().out


def test_test_thumos_script_smoke(monkeypatch, tmp_path, capsys):
    """tools/test_thumos.py: checkpoint in the reference's layout -> sliding windows over every test video -> device-side decode
    + soft-NMS (emulated by the oracle's restatements) -> detection json in the reference's layout."""
    abi_emu.install(monkeypatch)
    info, anno, cls, npy = make_golden.dataset_case_files(str(tmp_path / "data"), seed=2, n_videos=1)
    sd = O.synthetic_state_dict(O.OracleConfig(num_classes=20), loc_bias_shift=3.4657)
    os.makedirs(tmp_path / "ckpt_0")
    torch.save(sd, tmp_path / "ckpt_0" / "checkpoint-3.ckpt")
    os.symlink(tmp_path / "ckpt_0" / "checkpoint-3.ckpt", tmp_path / "ckpt_0" / "checkpoint-latest.ckpt")       # train.py:96-103
    yaml_text = make_golden.CONFIG_CASE_YAML
    for old, new in (("./data/open/split_{id:d}/test_info.csv", info), ("./data/open/split_{id:d}/test_anno.csv", anno),
                     ("./data/open/split_{id:d}/classes.txt", cls), ("./data/test_npy/", npy), ("backbone_model: ./weights/i3d.pt", "backbone_model: null"),
                     ("./ckpt/split_{id:d}/checkpoint-latest.ckpt", str(tmp_path / "ckpt_{id:d}" / "checkpoint-latest.ckpt")),
                     ("./output/split_{id:d}", str(tmp_path / "out_{id:d}"))):
        assert old in yaml_text, old
        yaml_text = yaml_text.replace(old, new)
    yaml_text = yaml_text.replace("num_classes: 16", "num_classes: 21").replace("conf_thresh: 0.01", "conf_thresh: 0.001").replace("top_k: 5000", "top_k: 50")
    cfg_path = tmp_path / "cfg.yaml"
    cfg_path.write_text(yaml_text)
    assert load_tool("test_thumos").main([str(cfg_path), "--open_set", "--split=0", "--device=cpu"]) == 0
    res = json.loads((tmp_path / "out_0" / "detection_results.json").read_text())
    assert res["version"] == "THUMOS14" and res["external_data"] == {} and len(res["results"]) == 1
    dets = next(iter(res["results"].values()))
    assert isinstance(dets, list)
    for d in dets:
        assert set(d) == {"label", "score", "segment", "uncertainty", "actionness"} and d["label"].startswith("Class") and len(d["segment"]) == 2
    assert "detections of 1 videos" in capsys.readouterr().out


def test_train_synthetic_tool_smoke(monkeypatch, tmp_path, capsys):
    """tools/train_synthetic.py — the first training run of the next GPU call — through the emulation: synthetic videos, window
    index, loader threads, train_loop.fit with the SSL pass, the IBM switch and the capture bookkeeping (eager here)."""
    abi_emu.install(monkeypatch)
    rc = load_tool("train_synthetic").main(["--videos", "2", "--epochs", "2", "--batch", "1", "--ibm-start", "2", "--steps-per-epoch", "1",
                                            "--no-graph", "--device", "cpu", "--loader-threads", "2", "--out", str(tmp_path / "out")])
    text = capsys.readouterr().out
    assert rc == 0 and "TRAIN_SYNTHETIC OK" in text and "windows from 2 videos" in text and "Epoch-2 Train Loss" in text
