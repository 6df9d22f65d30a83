"""The ActivityNet-1.3 training set as the hot path consumes it  (BASELINE configs[3]; AFSD/common/anet_dataset.py).

File formats are the reference's: one json `{video: {subset, frame_num, annotations: [{start_frame, end_frame, label_id}]}}`
(anet_dataset.py:32-40) and one `<video>.npy` per video, uint8 [frame_num <= 768, 112, 112, 3] (every video is resampled to
768 frames by the reference's preprocessing).  As in opental_b200/dataset.py a sample is the uint8 window as stored plus the
crop / mirror decision and the cut-paste frame map — the ingest kernel does the rest on the device — and the random draws
follow the reference's order (RandomCrop row, column; flip; cut-paste choices).

Reference behaviours kept (SURVEY App. D7 and anet_dataset.py):
  * one window per video at offset 0 (:66), `frame_num = min(frame_num, clip_length)`;
  * the action / start / end maps hold the CLASS ID, not 1, as BCE targets (:86-92) — `anet/train.py:134-143` trains on them;
  * cut-paste threshold `th = int(floor(shortest annotation) / 4)` (:104, :226), actions qualify with length >= 2*th, a paste
    that runs past the clip means "not augmented" (opental_b200.augment, variant="anet").
One deviation: a video shorter than the clip is padded by the reference with the value 127.5, i.e. exactly 0 after
normalisation (:236-239); uint8 frames cannot hold 127.5, so the pad here is 128 (+0.0039 after normalisation).  The
reference's preprocessing produces 768-frame videos, for which no padding happens."""
from __future__ import annotations

import json
import math
import os
import random as _random

import numpy as np

from . import augment
from .windows import annos_transform


def get_video_info(video_info_path: str, subset: str = "training") -> dict:
    """anet_dataset.py:32-40."""
    with open(video_info_path) as fh:
        data = json.load(fh)
    return {name: v for name, v in data.items() if v["subset"] == subset}


def split_videos(video_info: dict, clip_length: int, video_dir: str, binary_class: bool = False):
    """(training_list, min_anno_dict): anet_dataset.py:43-105.  Videos without an .npy file or without a valid annotation are
    skipped; maps `action`, `start`, `end` [clip_length] carry the class id."""
    training_list, min_anno_dict = [], {}
    for name, info in video_info.items():
        if not os.path.exists(os.path.join(video_dir, name + ".npy")):
            continue
        frame_num = min(info["frame_num"], clip_length)
        annos = []
        for a in info["annotations"]:
            label = (1 if a["label_id"] > 0 else 0) if binary_class else a["label_id"]
            if a["end_frame"] <= a["start_frame"]:
                continue
            annos.append([a["start_frame"], a["end_frame"], label])
        if not annos:
            continue
        min_anno = min(clip_length, min(x[1] - x[0] for x in annos))
        start, end, action = np.zeros([clip_length]), np.zeros([clip_length]), np.zeros([clip_length])
        for s, e, cid in annos:
            d = max((e - s) / 10.0, 2.0)
            clip = lambda v: int(np.clip(int(round(v)), 0, clip_length - 1))   # noqa: E731
            action[clip(s): clip(e) + 1] = cid
            start[clip(s - d / 2): clip(s + d / 2) + 1] = cid
            end[clip(e - d / 2): clip(e + d / 2) + 1] = cid
        training_list.append(dict(video_name=name, offset=0, annos=[list(a) for a in annos], frame_num=frame_num,
                                  start=start, end=end, action=action))
        min_anno_dict[name] = math.floor(min_anno)
    return training_list, min_anno_dict


class AnetWindows:
    """`ANET_Dataset` (anet_dataset.py:127-257) as an index of uint8 windows; `sample(idx, rng)` returns the dict of
    opental_b200.dataset.ThumosWindows.sample with scores float32 [3,clip_length] = (action, start, end)."""

    PAD_VALUE = 128      # the reference pads with 127.5 (see the module docstring)

    def __init__(self, video_info_path: str, video_dir: str, clip_length: int = 768, crop_size: int = 96, training: bool = True,
                 binary_class: bool = False):
        info = get_video_info(video_info_path, "training" if training else "validation")
        self.training_list, self.th = split_videos(info, clip_length, video_dir, binary_class)
        self.video_dir, self.clip_length, self.crop_size, self.training = video_dir, clip_length, crop_size, training

    def __len__(self) -> int:
        return len(self.training_list)

    def sample(self, idx: int, rng=_random) -> dict:
        w = self.training_list[idx]
        L, cs = self.clip_length, self.crop_size
        th = int(self.th[w["video_name"]] / 4)
        data = np.load(os.path.join(self.video_dir, w["video_name"] + ".npy"), mmap_mode="r")
        frames = np.asarray(data[w["offset"]: min(w["offset"] + L, w["frame_num"])])
        if frames.shape[0] < L:
            frames = np.concatenate([frames, np.full([L - frames.shape[0], *frames.shape[1:]], self.PAD_VALUE, frames.dtype)], 0)
        H, W = frames.shape[1], frames.shape[2]
        if self.training:
            i = rng.randint(0, H - cs) if H != cs else 0
            j = rng.randint(0, W - cs) if W != cs else 0
            flip = int(rng.random() < 0.5)
        else:
            i, j, flip = int(np.round((H - cs) / 2.0)), int(np.round((W - cs) / 2.0)), 0
        fmap, ssl_annos, flag = augment.cut_paste(w["annos"], th, L, 1, rng=rng, variant="anet")
        return dict(frames=frames, crop=(i, j, flip),
                    target=np.asarray(annos_transform(w["annos"], L), dtype=np.float32),
                    scores=np.stack([w["action"], w["start"], w["end"]]).astype(np.float32),
                    frame_map=fmap, ssl_target=np.asarray([a[:2] for a in ssl_annos], dtype=np.float32), flag=bool(flag),
                    video_name=w["video_name"], offset=w["offset"])
