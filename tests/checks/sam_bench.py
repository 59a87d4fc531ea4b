"""SAM-2.1 Hiera-L stage timing on one GPU (development aid): set_image / predict(256 prompts) / generate, CUDA events,
plus the per-kernel-class breakdown of ovo_profile_begin/report.
    python tests/checks/sam_bench.py [--tiny] [--iters 10]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gen_golden as GG  # noqa: E402  (image generator only)
from ovo_b200 import _lib  # noqa: E402
from ovo_b200.sam import Sam2  # noqa: E402
from ovo_b200.sam_config import SamConfig, random_state_dict, tiny_sam_config  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--points", type=int, default=16)
    ap.add_argument("--batch", type=int, default=4, help="frames per batched trunk pass")
    args = ap.parse_args()
    cfg = tiny_sam_config() if args.tiny else SamConfig()
    sam = Sam2(cfg, random_state_dict(cfg, seed=0), max_h=480, max_w=640, max_prompts=args.points ** 2, max_batch=args.batch)
    img = torch.from_numpy(GG.sam_image()).cuda()
    n = args.points
    from oracle import sam as OS
    pts = torch.from_numpy(OS.amg_points(n, 480, 640, cfg.image_size)).cuda()
    prm = sam.amg_params(points_per_side=n, pred_iou_thresh=0.45, stability_score_thresh=0.4, box_nms_thresh=0.7, nms_score_th=0.2)
    out = {"config": "tiny" if args.tiny else "hiera_l", "prompts": n * n}
    out["set_image_ms"] = timed(lambda: sam.set_image(img), args.iters)
    out["predict_ms"] = timed(lambda: sam.predict(pts), args.iters)
    out["generate_ms"] = timed(lambda: sam.generate(img, prm), args.iters)
    seg, maps = sam.generate(img, prm)
    out["masks"] = int(maps.shape[0])
    if args.batch > 1:
        imgs = img[None].repeat(args.batch, 1, 1, 1).contiguous()
        out["set_images_ms_per_frame"] = timed(lambda: sam.set_images(imgs), args.iters) / args.batch
        out["generate_batch_ms_per_frame"] = timed(lambda: sam.generate_batch(imgs, prm), args.iters) / args.batch
        _lib.profile_begin()
        sam.set_images(imgs)
        rep = _lib.profile_report()
        out["set_images_classes_per_frame"] = {k: {"ms": round(v["ms"] / args.batch, 4), "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1), "launches": v["launches"]}
                                               for k, v in rep.items() if v["launches"]}
        sam.set_image(img)
    import time
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5):
        sam.generate(img, prm)
    torch.cuda.synchronize(); out["generate_wall_ms"] = (time.time() - t0) / 5 * 1e3
    low, iou = sam.predict(pts)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5):
        sam.postprocess(low, iou, 480, 640, prm)
    torch.cuda.synchronize(); out["postprocess_wall_ms"] = (time.time() - t0) / 5 * 1e3
    for name, fn in (("set_image", lambda: sam.set_image(img)), ("predict", lambda: sam.predict(pts)), ("generate", lambda: sam.generate(img, prm))):
        _lib.profile_begin()
        fn()
        rep = _lib.profile_report()
        out[name + "_classes"] = {k: {"ms": round(v["ms"], 4), "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1), "launches": v["launches"]}
                                  for k, v in rep.items() if v["launches"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
