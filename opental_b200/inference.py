"""Sliding-window inference with device-side post-processing.

Replaces the per-video body of `test()` in AFSD/thumos14/test.py:203-256: clip offsets (`get_offsets`, :48-56),
`prepare_clip` (:67-76), `net(clip)`, `decode_predictions` (:112-140), `filtering` (:143-162) and per-class
`softnms_v2` (AFSD/common/segment_utils.py:128-162 — a Python loop on the CPU in the reference).  Everything up to the
final list of detections stays on the GPU: the windows of a video are one batch, decoding is one kernel, soft-NMS is
one CTA per class; one device->host copy returns the kept rows.
"""
from __future__ import annotations

import torch

from . import ops


def clip_offsets(sample_count: int, clip_length: int, stride: int) -> list[int]:
    """get_offsets (test.py:48-56)."""
    if sample_count < clip_length:
        return [0]
    offs = list(range(0, sample_count - clip_length + 1, stride))
    if (sample_count - clip_length) % stride:
        offs.append(sample_count - clip_length)
    return offs


@torch.no_grad()
def detect_video(net, frames: torch.Tensor, sample_fps: float, *, clip_length: int = 256, stride: int = 128,
                 conf_thresh: float = 0.01, top_k: int = 5000, nms_sigma: float = 0.5, batch: int = 8):
    """frames: normalised fp32 video [3, T, 96, 96] on the device (prepare_data + the normalisation of prepare_clip).
    Returns a dict class_index -> tensor [n, 5] of (start s, end s, score, uncertainty, actionness), on the host.
    Class indices are the model's (0..K-1 with the open-set head, test.py:208)."""
    assert frames.is_cuda and frames.dim() == 4
    T = frames.shape[1]
    offs = clip_offsets(T, clip_length, stride)
    segs, scores, uncts, acts = [], [], [], []
    for i in range(0, len(offs), batch):
        chunk = offs[i:i + batch]
        clips = torch.zeros(len(chunk), 3, clip_length, frames.shape[2], frames.shape[3], device=frames.device)
        for j, o in enumerate(chunk):                       # prepare_clip: zero padding (in normalised space) of a short tail
            n = min(clip_length, T - o)
            clips[j, :, :n] = frames[:, o:o + n]
        out = net(clips)
        s, sc, u, a = ops.decode_scores(out, torch.tensor(chunk, dtype=torch.float32), clip_length, sample_fps)
        segs.append(s); scores.append(sc); uncts.append(u); acts.append(a)
    seg = torch.cat(segs, 0).reshape(-1, 2)                             # [W*P, 2]
    score = torch.cat(scores, 0).permute(1, 0, 2).reshape(scores[0].shape[1], -1)      # [K, W*P]
    unct = torch.cat(uncts, 0).reshape(-1)
    act = torch.cat(acts, 0).reshape(-1)
    # filtering (test.py:143-162): below-threshold or low-actionness candidates never enter the NMS
    ok = score > conf_thresh
    if out.get("act") is not None:
        ok = ok & (act > 0.5).unsqueeze(0)
    else:
        # closed set (no open-set head): class 0 is the background — the reference never filters / suppresses / reports it
        # (`class_range = range(1, num_classes)`, test.py:208)
        ok[0] = False
    score = torch.where(ok, score, torch.zeros_like(score))
    decayed, keep, _ = ops.softnms(seg, score, sigma=nms_sigma, top_k=top_k, score_threshold=0.001)
    keep = keep & ok
    keep_c, decayed_c, seg_c, unct_c, act_c = keep.cpu(), decayed.cpu(), seg.cpu(), unct.cpu(), act.cpu()
    result = {}
    for cl in range(score.shape[0]):
        m = keep_c[cl]
        if m.any():
            result[cl] = torch.cat([seg_c[m], decayed_c[cl][m, None], unct_c[m, None], act_c[m, None]], -1)
    return result


def to_proposal_list(result: dict, idx_to_class: dict, *, os_head: bool = True, use_edl: bool = True) -> list[dict]:
    """The per-video list `get_video_detections` returns (test.py:182-200) from `detect_video`'s result: one dict per kept
    detection with the keys the evaluation reads (`label`, `score`, `segment`, `uncertainty`, `actionness`), classes in
    increasing index order, rows in descending score order (the order soft-NMS selects them in).  `idx_to_class` maps
    1..K to names (`get_class_index_map`, thumos_dataset.py:13-21); with the open-set head the model's class c is c + 1."""
    out = []
    for cl in sorted(result):
        rows = result[cl]
        name = idx_to_class[cl + 1 if os_head else cl]
        order = torch.argsort(rows[:, 2], descending=True, stable=True)
        for r in rows[order].tolist():
            if r[2] <= 0:
                continue
            out.append({"label": name, "score": float(r[2]), "segment": [float(r[0]), float(r[1])],
                        "uncertainty": float(r[3]) if use_edl else 0.0, "actionness": float(r[4]) if os_head else 0.0})
    return out


def results_json(per_video: dict, version: str = "THUMOS14") -> dict:
    """The file `test()` writes (test.py:253-255): {'version', 'results': {video: proposal list}, 'external_data'}."""
    return {"version": version, "results": dict(per_video), "external_data": {}}
