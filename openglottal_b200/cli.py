"""``openglottal run VIDEO --pipeline unet-only`` on the B200-native path.

Same flags, same ``features.json`` keys (including ``_area``) and the same exit code 1 on a
silent result as /root/reference/openglottal/cli.py:9-41,58-66,90-103. The YOLO/VFT pipelines
are out of scope here and stay with the reference package.
"""
from __future__ import annotations

import argparse
import sys


def main(argv: list[str] | None = None) -> None:
    parser = argparse.ArgumentParser(
        prog="openglottal",
        description="Glottal area segmentation from high-speed videoendoscopy (B200-native "
                    "unet-only pipeline).",
    )
    sub = parser.add_subparsers(dest="command", required=True)
    run_p = sub.add_parser("run", help="Run inference on a video file.")
    run_p.add_argument("video", help="Path to input .avi / .mp4 video.")
    run_p.add_argument("--yolo-weights", help="Unused here (YOLO pipelines stay with the reference).")
    run_p.add_argument("--unet-weights", help="Path to U-Net .pt weights (state dict).")
    run_p.add_argument("--pipeline", choices=["vft", "guided-vft", "unet", "unet-only"],
                       default="unet-only",
                       help="Only unet-only (no YOLO gate) is implemented natively.")
    run_p.add_argument("--output", "-o", default="results", help="Output directory.")
    run_p.add_argument("--device", default="cuda", help="Torch device (cuda / cuda:N).")
    run_p.add_argument("--precision", choices=["bf16", "fp16", "fp32"], default="bf16",
                       help="Operand type of the native kernels (not a reference flag): bf16 "
                            "tensor cores (default), fp16 tensor cores, or the fp32 validation mode.")
    args = parser.parse_args(argv)
    if args.command == "run":
        _cmd_run(parser, args)


def _cmd_run(parser: argparse.ArgumentParser, args: argparse.Namespace) -> None:
    import json
    import os

    import torch

    from .features import extract_features_unet
    from .unet import UNet

    if args.pipeline != "unet-only":
        parser.error(f"pipeline {args.pipeline!r} needs the YOLO detector / motion trackers, "
                     "which openglottal_b200 leaves to the reference package; use unet-only.")
    if not args.unet_weights:
        parser.error("--unet-weights is required for the unet-only pipeline.")
    device = torch.device(args.device)
    if device.type != "cuda":
        parser.error("openglottal_b200 runs on CUDA (B200) only; there is no CPU fallback.")

    model = UNet(1, 1, (32, 64, 128, 256)).to(device)
    model.load_state_dict(torch.load(args.unet_weights, map_location=device, weights_only=True))
    model.eval()
    model.precision = args.precision
    feats = extract_features_unet(args.video, None, model, device)

    if feats is None:
        print("No glottis detected — check your weights or input video.")
        sys.exit(1)

    os.makedirs(args.output, exist_ok=True)
    out_path = os.path.join(args.output, "features.json")
    save = {k: v.tolist() if hasattr(v, "tolist") else v for k, v in feats.items()}
    with open(out_path, "w") as f:
        json.dump(save, f, indent=2)
    print(f"Features saved to {out_path}")
    for k, v in feats.items():
        if not k.startswith("_"):
            print(f"  {k}: {v:.4f}" if isinstance(v, float) else f"  {k}: {v}")


if __name__ == "__main__":
    main()
