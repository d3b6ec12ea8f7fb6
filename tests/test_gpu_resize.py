"""GPU parity of the reference-resize mode (kernel K9): the two
cv2.resize calls of /root/reference/openglottal/utils.py:234,238-240 as CUDA kernels.

u8 squash: bit-exact with cv2 (integer work). Probability resize + threshold: bit-exact with the
NumPy restatement of OpenCV's f32 arithmetic except where the device's expf and NumPy's exp differ
in the last bit at a pixel sitting exactly on the threshold (none expected, <= 2 allowed)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [(512, 256), (256, 512), (512, 512), (96, 128), (300, 200), (1024, 1024), (128, 128),
         (257, 255), (480, 640), (64, 64), (256, 256), (320, 256), (256, 1000), (17, 33)]


def test_resize_u8_bit_exact_with_cv2(lib):
    import cv2
    import openglottal_b200 as ogl
    from oracle import resize_oracle as ro

    rng = np.random.default_rng(0)
    for hgt, wid in SIZES:
        frames = rng.integers(0, 256, (3, hgt, wid), dtype=np.uint8)
        got = ogl.resize_u8_linear(torch.from_numpy(frames).cuda()).cpu().numpy()
        assert got.shape == (3, 256, 256)
        for i in range(3):
            want = cv2.resize(frames[i], (256, 256), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(got[i], want), (hgt, wid, i)
            assert np.array_equal(got[i], ro.resize_u8_linear(frames[i], 256, 256)), (hgt, wid, i)


def test_prob_resize_mask_matches_restated_opencv(lib):
    import openglottal_b200 as ogl
    from oracle import resize_oracle as ro

    rng = np.random.default_rng(1)
    for hgt, wid in SIZES:
        logits = (rng.standard_normal((2, 256, 256)) * 3).astype(np.float32)
        # smooth blobs too: large connected regions like real masks
        yy, xx = np.mgrid[0:256, 0:256]
        logits[1] = (6.0 - 0.002 * ((yy - 120.0) ** 2 + 2 * (xx - 130.0) ** 2)).astype(np.float32)
        for thr in (0.5, 0.3):
            mask, area = ogl.prob_resize_mask(torch.from_numpy(logits).cuda(), hgt, wid, thr)
            mask, area = mask.cpu().numpy(), area.cpu().numpy()
            assert mask.shape == (2, hgt, wid) and set(np.unique(mask)) <= {0, 255}
            assert np.array_equal(area, (mask > 0).reshape(2, -1).sum(1))      # popcount bit-exact
            for i in range(2):
                want = ro.segment_frame_restated(logits[i], hgt, wid, thr)
                assert (mask[i] != want).sum() <= 2, (hgt, wid, thr, i)
    _, area_only = ogl.prob_resize_mask(torch.from_numpy(logits).cuda(), 512, 256, want_mask=False)
    m2, a2 = ogl.prob_resize_mask(torch.from_numpy(logits).cuda(), 512, 256)
    assert torch.equal(area_only, a2)


def test_reference_resize_fp32_masks_equal_oracle_512x256(lib, trained_sd):
    """The whole A2 row in the fp32 validation mode on BAGLS-shaped 512(H) x 256(W) frames against
    the oracle's unet_segment_frame (utils.py:218-241, cv2 resizes on the host): the same masks up to
    pixels whose probability is within the fp32 noise (1e-5) of the threshold."""
    import openglottal_b200 as ogl
    from oracle import synth, unet_oracle as uo

    clip, _ = synth.glottis_clip(4, 512, 256, seed=71, period=5.0)
    m = ogl.UNet().to("cuda")
    m.load_state_dict(trained_sd)
    m.eval()
    m.precision = "fp32"
    area, masks = ogl.segment_frames_reference_resize(torch.from_numpy(clip).cuda(), m)
    masks = masks.cpu().numpy()
    for i, f in enumerate(clip):
        want = uo.segment_frame(trained_sd, f)
        diff = int((masks[i] != want).sum())
        print("frame", i, "area", int(area[i]), "pixels differing from the oracle", diff)
        assert diff <= 2, i
        single = ogl.unet_segment_frame(f, m, torch.device("cuda"))
        assert np.array_equal(single, masks[i])
    assert np.array_equal(area.cpu().numpy(), (masks > 0).reshape(4, -1).sum(1))


def test_reference_resize_bf16_dice_512x256_and_odd_sizes(lib, native_model, trained_sd):
    """bf16 product path through the resize mode: Dice >= 0.999 against the oracle over a batch of
    512 x 256 frames (north_star's mask bar), per-frame areas within 0.5 % (or 2 px), and sizes that
    are not multiples of 16 (the reference accepts any frame size)."""
    import openglottal_b200 as ogl
    from oracle import synth, unet_oracle as uo

    clip, _ = synth.glottis_clip(8, 512, 256, seed=72, period=6.0)
    area, masks = ogl.masks_for_clip(clip, native_model, want_masks=True)
    masks, area = masks.cpu().numpy(), area.cpu().numpy()
    want = np.stack([uo.segment_frame(trained_sd, f) for f in clip])
    d = ogl.dice(masks, want)
    per = [int((a != b).sum()) for a, b in zip(masks, want)]
    print("512x256 bf16 resize mode: dice", d, "differing pixels per frame", per)
    assert d >= 0.999
    want_area = (want > 0).reshape(8, -1).sum(1)
    assert (np.abs(area - want_area) <= np.maximum(2, 0.005 * want_area)).all()
    odd, _ = synth.glottis_clip(2, 304, 200, seed=73, period=4.0)
    odd = np.ascontiguousarray(odd[:, :300, :])           # 300 x 200: not a multiple of 16
    a2, m2 = ogl.masks_for_clip(odd, native_model, want_masks=True)
    assert m2.shape == (2, 300, 200)
    w2 = np.stack([uo.segment_frame(trained_sd, f) for f in odd])
    assert ogl.dice(m2.cpu().numpy(), w2) >= 0.995
