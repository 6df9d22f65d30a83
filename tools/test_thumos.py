"""THUMOS14 inference with the reference's command line  (AFSD/thumos14/test.py:203-256): sliding windows over every test
video, decoding + filtering + per-class soft-NMS on the device (opental_b200/inference.py), detections written in the
reference's json layout so that AFSD/thumos14/eval*.py read them unchanged.

    python tools/test_thumos.py configs/thumos14_opental_final.yaml --open_set --split=0 [--output_json detection_results.json]

RGB only (`--fusion` needs the optical-flow twin model, SURVEY §2: out of scope).  Not exercised by the GPU test-suite (it
needs the dataset and a checkpoint); its parts are tests/test_infer_gpu.py and the config / dataset CPU tests."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from opental_b200 import config as C, dataset as D, inference  # noqa: E402
from opental_b200.bdnet import BDNet  # noqa: E402


def main(argv=None) -> int:
    parser = C.build_parser()
    parser.add_argument("--device", type=str, default="cuda", help="'cuda' (the product has no CPU path; other values are for the test harness)")
    args = parser.parse_args(argv)
    cfg = C.get_config(argv, parser)
    if cfg["testing"].get("fusion"):
        raise NotImplementedError("--fusion (RGB + optical-flow late fusion) is outside the OpenTAL hot path")
    te, ds, model = cfg["testing"], cfg["dataset"]["testing"], cfg["model"]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if args.device == "cuda" else torch.device(args.device)
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    os_head, use_edl = bool(model.get("os_head", False)), bool(model.get("use_edl", False))
    net = BDNet.from_config(cfg, training=False, use_edl=use_edl, use_rpl=False, frame_num=ds["clip_length"]).to(dev)
    ckpt = te["checkpoint_path"]
    net.load_state_dict(torch.load(os.path.realpath(ckpt), map_location="cpu"))          # get_path: follows checkpoint-latest
    net.eval()
    infos = D.get_video_info(ds["video_info_path"])
    _, idx_to_class = D.get_class_index_map(cfg["dataset"]["class_info_path"])
    crop = ds["crop_size"]
    results = {}
    for name, info in infos.items():
        video = np.load(os.path.join(ds["video_data_path"], name + ".npy"), mmap_mode="r")     # uint8 [T,112,112,3]
        i, j = int(np.round((video.shape[1] - crop) / 2.0)), int(np.round((video.shape[2] - crop) / 2.0))   # CenterCrop
        px = torch.from_numpy(np.ascontiguousarray(video[:, i:i + crop, j:j + crop, :])).to(dev)
        frames = (px.permute(3, 0, 1, 2).float() / 255.0) * 2.0 - 1.0                    # prepare_clip's normalisation
        res = inference.detect_video(net, frames, float(info["sample_fps"]), clip_length=ds["clip_length"], stride=ds["clip_stride"],
                                     conf_thresh=te["conf_thresh"], top_k=te["top_k"], nms_sigma=te["nms_sigma"])
        results[name] = inference.to_proposal_list(res, idx_to_class, os_head=os_head, use_edl=use_edl)
    os.makedirs(te["output_path"], exist_ok=True)
    out = os.path.join(te["output_path"], te["output_json"])
    with open(out, "w") as fh:
        json.dump(inference.results_json(results), fh)
    print(f"{sum(len(v) for v in results.values())} detections of {len(results)} videos -> {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
