"""`openglottal run VIDEO --pipeline unet-only` end to end on the GPU (SURVEY.md section 8 row A6,
/root/reference/openglottal/cli.py:58-66,90-103): features.json against the reference's own output
for the same clip and weights (tests/golden/pipeline.json, written by the unmodified reference),
and exit code 1 with the reference's message when nothing is segmented."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
KEYS = ["area_mean", "area_std", "area_range", "open_quotient", "f0", "periodicity", "cv", "_area"]


def test_cli_run_writes_reference_features_json(lib, calibrated_sd, tmp_path, capsys):
    """The golden was produced with the variance-calibrated RANDOM weights (logits of O(10) that
    amplify bf16 rounding), so the network runs in the fp32 validation mode: this test is about the
    CLI and the pipeline around the network (decode, gray, resize mode for the 64x64 clip, area,
    features, JSON). The CPU oracle itself reproduces the golden areas within 2 px
    (tests/test_oracle_golden.py), the same bar holds here."""
    from openglottal_b200 import cli

    weights = tmp_path / "openglottal_unet.pt"
    torch.save(calibrated_sd, weights)                       # scripts/train_unet.py:207 format
    out = tmp_path / "results"
    cli.main(["run", str(GOLDEN / "pipeline_clip.avi"), "--unet-weights", str(weights),
              "--pipeline", "unet-only", "--output", str(out), "--device", "cuda",
              "--precision", "fp32"])
    printed = capsys.readouterr().out
    assert f"Features saved to {out / 'features.json'}" in printed
    got = json.loads((out / "features.json").read_text())
    want = json.loads((GOLDEN / "pipeline.json").read_text())
    assert list(got.keys()) == KEYS == list(want.keys())
    assert isinstance(got["_area"], list) and len(got["_area"]) == len(want["_area"]) == 30
    assert all(isinstance(v, float) for v in got["_area"])
    err = np.abs(np.array(got["_area"]) - np.array(want["_area"]))
    print("cli: max |area - reference|", err.max())
    assert err.max() <= 2
    assert got["f0"] == want["f0"]
    assert got["area_mean"] == pytest.approx(want["area_mean"], rel=1e-3)
    assert got["open_quotient"] == want["open_quotient"]
    # the scalars in the file are the reference's _kinematic_features of the area in the file
    from oracle.features_oracle import kinematic_features

    again = kinematic_features(got["_area"], exact_correlate=True)
    for k in ("area_mean", "area_std", "area_range", "open_quotient", "periodicity", "cv"):
        assert got[k] == pytest.approx(float(again[k]), rel=1e-9, abs=1e-12), k
    assert got["f0"] == again["f0"]
    for k in ("area_mean", "area_std", "area_range", "open_quotient", "f0", "periodicity", "cv"):
        assert f"  {k}: " in printed                          # cli.py:100-103 prints every scalar


def test_cli_exits_1_when_nothing_is_segmented(lib, calibrated_sd, tmp_path, capsys):
    """cli.py:90-92: `features is None` (all areas zero, features.py:45-46) -> message + exit 1,
    and no features.json."""
    from openglottal_b200 import cli

    sd = {k: v.clone() for k, v in calibrated_sd.items()}
    sd["head.weight"].zero_()
    sd["head.bias"].fill_(-20.0)                             # every logit -20: empty masks
    weights = tmp_path / "silent.pt"
    torch.save(sd, weights)
    out = tmp_path / "results"
    with pytest.raises(SystemExit) as exc:
        cli.main(["run", str(GOLDEN / "pipeline_clip.avi"), "--unet-weights", str(weights),
                  "--output", str(out)])
    assert exc.value.code == 1
    assert "No glottis detected" in capsys.readouterr().out
    assert not (out / "features.json").exists()
