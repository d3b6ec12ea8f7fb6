"""BASELINE.json's full sizes (batch 512 at 256x256; 512x256 BAGLS-shaped frames), checked through
size-independent properties because the CPU oracle needs ~80 ms per frame:

  * a sample of the frames against the fp32 oracle (Dice >= 0.999, area within 0.5 %),
  * every frame's result is independent of its position in the batch and of the batch size
    (bit-exact: tiles never mix frames),
  * area == popcount(mask) for every frame (bit-exact),
  * the two kernel schedules of the full-resolution level agree,
  * two runs are bit-identical (no atomics on floating point, no uninitialised reads).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _clip(n, hgt=256, wid=256, seed=3, period=10.0):
    from oracle import synth

    return synth.glottis_clip(n, hgt, wid, seed=seed, period=period)[0]


@pytest.fixture(scope="module")
def batch512():
    base = _clip(64, seed=51, period=9.0)
    frames = np.concatenate([np.roll(base, 5 * r, axis=2) for r in range(8)])
    return torch.from_numpy(np.ascontiguousarray(frames)).cuda()


def test_batch512_properties(native_model, trained_sd, batch512):
    from oracle import unet_oracle as uo
    from openglottal_b200 import dice

    lg, mask, area = native_model.run(batch512, want_logits=True)
    assert mask.shape == (512, 256, 256) and area.shape == (512,)
    assert np.array_equal(area.cpu().numpy(), (mask > 0).reshape(512, -1).sum(1).cpu().numpy())
    assert set(torch.unique(mask).tolist()) <= {0, 255}
    # determinism
    lg2, mask2, area2 = native_model.run(batch512, want_logits=True)
    assert torch.equal(lg, lg2) and torch.equal(mask, mask2) and torch.equal(area, area2)
    # position / batch-size independence: a permuted batch and small ragged chunks
    perm = torch.randperm(512, generator=torch.Generator().manual_seed(0)).cuda()
    lg_p, mask_p, area_p = native_model.run(batch512[perm], want_logits=True)
    assert torch.equal(lg_p, lg[perm]) and torch.equal(mask_p, mask[perm]) and torch.equal(area_p, area[perm])
    native_model.max_batch = 37
    try:
        _, mask_c, area_c = native_model.run(batch512[:200])
    finally:
        native_model.max_batch = 512
    assert torch.equal(mask_c, mask[:200]) and torch.equal(area_c, area[:200])
    # a sample against the fp32 oracle
    idx = [0, 63, 64, 300, 511]
    ref_lg, ref_mask, ref_area = uo.batch_masks(trained_sd, batch512[idx].cpu().numpy())
    got_mask = mask[idx].cpu().numpy()
    assert dice(got_mask, ref_mask) >= 0.999
    rel = np.abs(area[idx].cpu().numpy() - ref_area) / np.maximum(ref_area, 1)
    assert rel.max() <= 0.005
    near = np.abs(ref_lg) < 1.0
    assert np.abs(lg[idx].cpu().numpy() - ref_lg)[near].max() <= 2e-2


def test_batch512_schedules_agree(native_model, batch512):
    _, mask, area = native_model.run(batch512)
    native_model.schedule = "direct"
    try:
        _, mask_d, area_d = native_model.run(batch512)
    finally:
        native_model.schedule = "s2d"
    diff = (mask != mask_d).sum().item()
    print("pixels that differ between schedules:", diff, "of", mask.numel())
    assert diff <= 2e-5 * mask.numel()
    assert (area - area_d).abs().max().item() <= 8


def test_bagls_shape_512x256_bf16(native_model, trained_sd):
    """BASELINE.json configs[2]: 512(H) x 256(W) frames, native mode (forward at 512x256)."""
    from oracle import unet_oracle as uo
    from openglottal_b200 import dice

    frames = _clip(6, 512, 256, seed=61, period=5.0)
    ref_lg, ref_mask, ref_area = uo.batch_masks(trained_sd, frames, chunk=3)
    lg, mask, area = native_model.run(torch.from_numpy(frames).cuda(), want_logits=True)
    mask, area = mask.cpu().numpy(), area.cpu().numpy()
    d = dice(mask, ref_mask)
    rel = np.abs(area - ref_area) / np.maximum(ref_area, 1)
    print("512x256: dice", d, "max rel area err", rel.max())
    assert d >= 0.999 and rel.max() <= 0.005
    assert np.array_equal(area, (mask > 0).reshape(6, -1).sum(1))
    near = np.abs(ref_lg) < 1.0
    assert np.abs(lg.cpu().numpy() - ref_lg)[near].max() <= 2e-2
    # odd tile counts in both directions (H/2 = 8 mod 16 at the s2d level)
    odd = _clip(3, 48, 80, seed=62)
    lg_o, m_o, a_o = native_model.run(torch.from_numpy(odd).cuda(), want_logits=True)
    bit = uo.folded_forward(trained_sd, uo.frames_to_input(odd), bf16=True)[:, 0].numpy()
    assert np.abs(lg_o.cpu().numpy() - bit).max() <= 3e-2


def test_million_frame_feature_step(lib):
    """BASELINE.json configs[4]: the feature step on a 10^6-sample waveform; round trip of the
    dominant frequency and exact integer statistics."""
    import openglottal_b200 as ogl

    n = 1_000_000
    t = np.arange(n)
    wave = np.floor(np.maximum(0, 1200 * np.sin(2 * np.pi * t / 16.0)) + 30).astype(np.int32)
    got = ogl.kinematic_features_device(torch.from_numpy(wave).cuda())
    assert got["f0"] == pytest.approx(1 / 16.0, rel=1e-12)
    assert float(got["area_mean"]) == pytest.approx(wave.astype(np.float64).mean(), rel=1e-12)
    assert float(got["area_range"]) == float(wave.max() - wave.min())
    assert got["periodicity"] > 0.99


def test_config1_pipeline_batch32_vs_oracle(native_model, trained_sd):
    """BASELINE.json configs[0] in miniature: the unet-only pipeline at batch 32 on a 256x256
    clip (240 frames, 12 periods) against the fp32 oracle run on the host cores: masks
    (Dice >= 0.999), area waveform (0.5 %), kinematic features (1e-3 relative; f0 exact)."""
    import openglottal_b200 as ogl
    from oracle import unet_oracle as uo
    from oracle.features_oracle import kinematic_features

    frames = _clip(240, seed=91, period=20.0)
    _, ref_mask, ref_area = uo.batch_masks(trained_sd, frames, chunk=16)
    area, masks = ogl.segment_clip(torch.from_numpy(frames), native_model, batch=32, want_masks=True)
    got_area = area.cpu().numpy()
    assert ogl.dice(masks.cpu().numpy(), ref_mask) >= 0.999
    rel = np.abs(got_area - ref_area) / np.maximum(ref_area, 1)
    assert rel.max() <= 0.005
    got = ogl.kinematic_features_device(area)
    want = kinematic_features(ref_area.astype(np.float64))
    print({k: (got[k], want[k]) for k in ("area_mean", "open_quotient", "f0", "periodicity")})
    for k in ("area_mean", "area_std", "area_range", "open_quotient", "periodicity", "cv"):
        assert abs(float(got[k]) - float(want[k])) <= 1e-3 * max(abs(float(want[k])), 1e-9) + 1e-9, k
    assert got["f0"] == want["f0"] == 0.05
