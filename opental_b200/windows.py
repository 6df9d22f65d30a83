"""Sliding-window index of the training set and the start / end score maps  (the tensors `Trainer.step`
consumes besides the clips).  Host-side, run once per dataset — numpy, like the reference's loader.

`split_videos` follows AFSD/common/thumos_dataset.py:69-130: windows of `clip_length` sampled frames every `stride`
frames (plus one flush with the end of the video), ground-truth segments kept when at least half of them lies inside the
window (clipped to [1, clip_length]), a window kept when it contains at least one segment completely; per window the
boundary score maps of `boundary_score_maps`; per video the cut-paste threshold `th` = ceil of the shortest kept segment.
Pinned to the reference's own function on synthetic annotation tables (oracle/make_golden.py --windows)."""
from __future__ import annotations

import math

import numpy as np


def annos_transform(annos, clip_length: int):
    """Frames -> fractions of the clip (thumos_dataset.py:58-66)."""
    return [[a[0] * 1.0 / clip_length, a[1] * 1.0 / clip_length, a[2]] for a in annos]


def boundary_score_maps(annos, clip_length: int) -> tuple[np.ndarray, np.ndarray]:
    """start / end maps [clip_length] (thumos_dataset.py:109-120): ones inside a band of width d = max(len / 10, 2)
    frames centred on each start / end boundary (python round = half-to-even, clipped to the clip)."""
    start, end = np.zeros([clip_length]), np.zeros([clip_length])
    for s, e, _ in annos:
        d = max((e - s) / 10.0, 2.0)
        for arr, c in ((start, s), (end, e)):
            lo = int(np.clip(int(round(c - d / 2.0)), 0, clip_length - 1))
            hi = int(np.clip(int(round(c + d / 2.0)), 0, clip_length - 1)) + 1
            arr[lo:hi] = 1
    return start, end


def split_videos(video_infos: dict, video_annos: dict, clip_length: int = 256, stride: int = 30):
    """(training_list, th).  video_infos[name]['sample_count'] = sampled frames of the video; video_annos[name] =
    [[start, end, label], ...] in sampled frames.  training_list entries: video_name, offset, annos (window-relative
    frames), start, end."""
    training_list, th = [], {}
    for name, annos in video_annos.items():
        min_anno = clip_length
        count = video_infos[name]["sample_count"]
        if count <= clip_length:
            offsets = [0]
            min_anno = min(min_anno, min(a[1] - a[0] for a in annos))
        else:
            offsets = list(range(0, count - clip_length + 1, stride))
            if (count - clip_length) % stride:
                offsets.append(count - clip_length)
        for offset in offsets:
            left, right = offset + 1, offset + clip_length
            cur, complete = [], False
            for a in annos:
                ioa = (min(right, a[1]) - max(left, a[0])) * 1.0 / (a[1] - a[0])
                complete |= ioa >= 1.0
                if ioa >= 0.5:
                    cur.append([max(a[0] - offset, 1), min(a[1] - offset, clip_length), a[2]])
            if cur:
                min_anno = min(min_anno, min(a[1] - a[0] for a in cur))
            if complete:
                start, end = boundary_score_maps(cur, clip_length)
                training_list.append(dict(video_name=name, offset=offset, annos=cur, start=start, end=end))
        th[name] = math.ceil(min_anno)
    return training_list, th
