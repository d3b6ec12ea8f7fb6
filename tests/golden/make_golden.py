"""Generates the committed golden fixtures by running the UNMODIFIED reference
(/root/reference, importable only in the build container) on seeded inputs.

    python tests/golden/make_golden.py

Outputs (all small, committed):
  features_kat.json        reference `_kinematic_features` on known waveforms
  unet_forward.npz         reference `UNet.forward` logits, calibrated_state(0), 2 frames 64x64
  segment_frame.npz        reference `unet_segment_frame` masks (256x256 identity path and a
                           96x128 frame through both cv2 resizes), bit-packed
  pipeline_clip.avi        30-frame 64x64 lossless (FFV1) synthetic clip
  pipeline.json            reference `extract_features_unet(clip, None, model, cpu)` output

The state dict is regenerated from its seed at test time (oracle.synth.calibrated_state uses only
a CPU torch.Generator), so no weights are committed.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
stub = types.ModuleType("ultralytics")   # openglottal/models/detector.py:6 imports it; YOLO unused
stub.YOLO = object
sys.modules.setdefault("ultralytics", stub)
sys.path.insert(0, "/root/reference")

import cv2  # noqa: E402
from openglottal.features import _kinematic_features, extract_features_unet  # noqa: E402
from openglottal.models.unet import UNet  # noqa: E402
from openglottal.utils import unet_segment_frame  # noqa: E402

from oracle import synth  # noqa: E402


def jsonable(d):
    if d is None:
        return None
    return {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in d.items()}


def features_kat():
    t500 = np.arange(500)
    cases = [
        [0.0] * 10,
        [5.0],
        [5.0, 7.0],
        [5.0, 7.0, 5.0],
        [100.0] * 64,
        np.floor(np.maximum(0, 300 * np.sin(2 * np.pi * t500 / 20))).tolist(),
        np.floor(200 + 150 * np.sin(2 * np.pi * 0.06 * t500 + 0.3)).tolist(),
        np.random.default_rng(1234).integers(0, 65537, 1000).astype(float).tolist(),
        np.floor(np.maximum(0, 800 * np.sin(2 * np.pi * np.arange(2000) / 12.5)) + 40).tolist(),
        (np.random.default_rng(5).random(301) * 1000).tolist(),          # non-integral, odd n
        np.floor(500 + 400 * np.cos(2 * np.pi * np.arange(49) / 7)).tolist(),   # n < 50 lags
    ]
    out = []
    for wave in cases:
        try:
            res = _kinematic_features(wave)
            if res is not None:
                res = {k: v for k, v in jsonable(res).items() if k != "_area"}
            out.append({"input": wave, "raises": False, "output": res})
        except ValueError as e:
            out.append({"input": wave, "raises": True, "error": str(e), "output": None})
    (HERE / "features_kat.json").write_text(json.dumps(out))


def main():
    features_kat()
    sd = synth.calibrated_state(0)
    model = UNet(1, 1, (32, 64, 128, 256))
    model.load_state_dict(sd)
    model.eval()

    frames, _ = synth.glottis_clip(2, 64, 64, seed=11, period=5.0)
    with torch.no_grad():
        logits = model(torch.from_numpy(frames.astype("float32") / 255.0).unsqueeze(1)).numpy()
    np.savez_compressed(HERE / "unet_forward.npz", frames=frames, logits=logits)

    f256 = synth.glottis_clip(1, 256, 256, seed=12)[0][0]
    f96 = synth.glottis_clip(1, 96, 128, seed=13)[0][0]
    m256 = unet_segment_frame(f256, model, torch.device("cpu"))
    m96 = unet_segment_frame(f96, model, torch.device("cpu"))
    np.savez_compressed(HERE / "segment_frame.npz", f256=f256, f96=f96,
                        m256=np.packbits(m256 > 0), m96=np.packbits(m96 > 0))

    clip, _ = synth.glottis_clip(30, 64, 64, seed=14, period=6.0)
    path = HERE / "pipeline_clip.avi"
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), 25.0, (64, 64))
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    cap = cv2.VideoCapture(str(path))
    back = []
    while True:
        ok, frm = cap.read()
        if not ok:
            break
        back.append(cv2.cvtColor(frm, cv2.COLOR_BGR2GRAY))
    cap.release()
    assert len(back) == 30 and np.array_equal(np.stack(back), clip), "codec is not lossless here"
    feats = extract_features_unet(str(path), None, model, torch.device("cpu"))
    (HERE / "pipeline.json").write_text(json.dumps(jsonable(feats)))
    print("golden written:", sorted(p.name for p in HERE.iterdir() if p.is_file()))


if __name__ == "__main__":
    main()
