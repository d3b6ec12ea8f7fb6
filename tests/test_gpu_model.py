"""Model-level parity of the CUDA path against the oracle (tests/ is the only place, with
smoke() and bench.py's cpu_baseline, that may touch oracle/).

Tolerances (BASELINE.json north_star): logits max-abs <= 2e-2 in bf16 on trained-like weights
(<= 1e-4 in the fp32 validation mode); mask Dice >= 0.999; area within 0.5 %; popcount of a
given mask bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _clip(n, hgt=256, wid=256, seed=3, period=10.0):
    from oracle import synth

    return synth.glottis_clip(n, hgt, wid, seed=seed, period=period)[0]


def test_state_dict_roundtrip(native_model, trained_sd):
    sd = native_model.state_dict()
    assert list(sd.keys())[:3] == ["downs.0.net.0.weight", "downs.0.net.1.weight", "downs.0.net.1.bias"]
    assert len(sd) == 118
    for k, v in trained_sd.items():
        assert torch.equal(sd[k].cpu(), v), k


def test_fp32_mode_matches_reference_forward(native_model, trained_sd):
    """fp32 validation mode: |logit - oracle| <= 1e-4."""
    from oracle import unet_oracle as uo

    frames = _clip(4)
    ref = uo.ref_forward(trained_sd, uo.frames_to_input(frames))[:, 0].numpy()
    native_model.precision = "fp32"
    try:
        lg, mask, area = native_model.run(torch.from_numpy(frames).cuda(), want_logits=True)
    finally:
        native_model.precision = "bf16"
    err = np.abs(lg.cpu().numpy() - ref).max()
    print("fp32-mode max|err|", err)
    assert err <= 1e-4
    ref_mask = (ref > 0)
    assert np.array_equal(mask.cpu().numpy() > 0, ref_mask) or (
        (mask.cpu().numpy() > 0) != ref_mask).sum() <= 2
    assert np.array_equal(area.cpu().numpy(), (mask.cpu().numpy() > 0).reshape(4, -1).sum(1))


def test_fp32_mode_calibrated_weights_512x256(lib, calibrated_sd):
    """Second architecture-size case (BAGLS-shaped 512x256), random calibrated weights:
    relative tolerance because logits are O(10)."""
    import openglottal_b200 as ogl
    from oracle import unet_oracle as uo

    m = ogl.UNet().to("cuda")
    m.load_state_dict(calibrated_sd)
    m.eval()
    m.precision = "fp32"
    frames = _clip(2, 512, 256)
    ref = uo.ref_forward(calibrated_sd, uo.frames_to_input(frames))[:, 0].numpy()
    lg, _, _ = m.run(torch.from_numpy(frames).cuda(), want_logits=True)
    err = np.abs(lg.cpu().numpy() - ref).max()
    print("fp32-mode calibrated max|err|", err, "logit rms", np.sqrt((ref ** 2).mean()))
    assert err <= 2e-4 * max(1.0, np.abs(ref).max())


def test_bf16_matches_bitmodel(native_model, trained_sd):
    """Kernel vs the CPU bit-model (same rounding points): tight tolerance, catches indexing
    bugs independently of bf16 noise."""
    from oracle import unet_oracle as uo

    frames = _clip(3)
    bit = uo.folded_forward(trained_sd, uo.frames_to_input(frames), bf16=True)[:, 0].numpy()
    lg, _, _ = native_model.run(torch.from_numpy(frames).cuda(), want_logits=True)
    err = np.abs(lg.cpu().numpy() - bit)
    print("bf16 vs bit-model max|err|", err.max(), "mean", err.mean())
    assert err.max() <= 3e-2 and err.mean() <= 2e-3


def test_direct_schedule_matches_bitmodel_and_s2d(eager_model, trained_sd):
    native_model = eager_model
    """The per-tap form of the full-resolution level (schedule "direct") against its own
    bit-model, and the two schedules against each other (masks equal up to boundary pixels)."""
    from oracle import unet_oracle as uo

    frames = _clip(3)
    dev = torch.from_numpy(frames).cuda()
    lg_s2d, m_s2d, a_s2d = native_model.run(dev, want_logits=True)
    native_model.schedule = "direct"
    native_model.compose_up = False      # every ConvTranspose2d its own launch, `up` rounded to bf16
    try:
        lg, m, a = native_model.run(dev, want_logits=True)
        assert _native_launches(native_model) == 22
    finally:
        native_model.schedule = "s2d"
        native_model.compose_up = True
    bit = uo.folded_forward(trained_sd, uo.frames_to_input(frames), bf16=True,
                            composed_level0=False, composed_up=False)[:, 0].numpy()
    err = np.abs(lg.cpu().numpy() - bit)
    print("direct vs bit-model max|err|", err.max(), "mean", err.mean())
    assert err.max() <= 3e-2 and err.mean() <= 2e-3
    d = (lg - lg_s2d).abs()
    print("direct vs s2d max|diff|", d.max().item(), "mean", d.mean().item())
    assert d.max().item() <= 5e-2 and d.mean().item() <= 3e-3
    assert (m != m_s2d).sum().item() <= 1e-4 * m.numel()
    assert np.array_equal(a.cpu().numpy(), (m.cpu().numpy() > 0).reshape(3, -1).sum(1))


def _native_launches(model):
    """Launches of the most recent EAGER forward (a forward replayed as a CUDA graph -- small
    batches, from the second call on -- does not pass through ogl_unet_forward)."""
    from openglottal_b200 import _native

    return _native.load().ogl_unet_launch_count(model._handle)


@pytest.fixture
def eager_model(native_model):
    native_model.use_graphs = False
    try:
        yield native_model
    finally:
        native_model.use_graphs = True


def test_composed_decoder_17_launches_and_agrees_with_separate_convT(eager_model, trained_sd):
    native_model = eager_model
    """Decoder levels 1-3 with every ConvTranspose2d composed into the conv after it (upcat_tc.cu,
    the default: 17 launches, no `up` tensor) against the round-1 schedule (20 launches: separate
    transposed convs, `up` rounded to bf16) and against the bit-model of each; incl. partial tiles,
    a 512x256 batch and a batch with several tiles per SM (CTA pairs)."""
    from oracle import unet_oracle as uo

    for shape in ((3, 256, 256), (2, 48, 80), (1, 16, 16), (2, 512, 256), (40, 256, 256)):
        frames = _clip(*shape)
        dev = torch.from_numpy(frames).cuda()
        lg_c, m_c, a_c = native_model.run(dev, want_logits=True)
        assert _native_launches(native_model) == 17
        native_model.compose_up = False
        try:
            lg_s, m_s, a_s = native_model.run(dev, want_logits=True)
            assert _native_launches(native_model) == 20
        finally:
            native_model.compose_up = True
        d = (lg_c - lg_s).abs()
        print(shape, "composed vs separate convT: max|dz|", d.max().item(), "mean", d.mean().item())
        assert d.max().item() <= 5e-2 and d.mean().item() <= 3e-3, shape
        assert (m_c != m_s).sum().item() <= 1e-4 * m_c.numel() + 2, shape
        assert torch.equal(a_c.cpu(), (m_c > 0).flatten(1).sum(1).to(torch.int32).cpu()), shape
        if shape[0] <= 3:
            x = uo.frames_to_input(frames)
            for lg, comp in ((lg_c, True), (lg_s, False)):
                bit = uo.folded_forward(trained_sd, x, bf16=True, composed_up=comp)[:, 0].numpy()
                err = np.abs(lg.cpu().numpy() - bit)
                print(shape, "composed" if comp else "separate", "vs its bit-model: max", err.max(), "mean", err.mean())
                assert err.max() <= 3e-2 and err.mean() <= 2e-3, (shape, comp)


def test_cta_pairs_bit_identical(native_model):
    """tcgen05 cta_group::2 for the Cout >= 64 conv layers: same logits, masks and areas, bit for
    bit, on a small batch (pairs forced) and on a batch with a tile per SM (the default rule)."""
    frames = torch.from_numpy(_clip(5)).cuda()
    big = torch.from_numpy(np.concatenate([_clip(64, seed=s) for s in (71, 72, 73, 74, 75)])).cuda()
    default = native_model.cta_pairs
    try:
        native_model.cta_pairs = 1
        ref = native_model.run(frames, want_logits=True)
        ref_big = native_model.run(big)
        native_model.cta_pairs = 3
        got = native_model.run(frames, want_logits=True)
        native_model.cta_pairs = 2
        got_big = native_model.run(big)
    finally:
        native_model.cta_pairs = default
    assert all(torch.equal(a, b) for a, b in zip(ref, got))
    assert torch.equal(ref_big[1], got_big[1]) and torch.equal(ref_big[2], got_big[2])


def test_fused_stem_bit_identical(native_model):
    """The stem computed inside the downs.0.net.3 kernel (u8 input) equals the separate stem
    kernel bit for bit -- same fp32 FMA order, same rounding point -- incl. partial tiles."""
    default = native_model.fuse_stem
    for shape in ((5, 256, 256), (3, 48, 80), (2, 512, 256), (1, 16, 16)):
        frames = torch.from_numpy(_clip(*shape)).cuda()
        try:
            native_model.fuse_stem = 0
            ref = native_model.run(frames, want_logits=True)
            native_model.fuse_stem = 1
            got = native_model.run(frames, want_logits=True)
        finally:
            native_model.fuse_stem = default
        assert all(torch.equal(a, b) for a, b in zip(ref, got)), shape


def test_tensor_core_stem_matches_fp32_stem(native_model):
    """The stem as a GEMM on the tensor cores (u8 taps exact in bf16, weights and bias split
    hi + lo in bf16, fp32 accumulation) against the fp32 CUDA-core stem: the stem outputs differ
    by ~2^-17 relative before their rounding to bf16, so a few of them round the other way and
    the logits move by a fraction of the bf16 noise; masks and areas almost everywhere equal.
    Incl. partial tiles, a single 16x16 frame and a batch with several tiles per SM."""
    default = native_model.fuse_stem
    shapes = ((5, 256, 256), (3, 48, 80), (2, 512, 256), (1, 16, 16), (40, 256, 256))
    g = torch.Generator().manual_seed(5)
    for shape in shapes + ("noise",):
        if shape == "noise":    # uniform random bytes: every tap large, no flat regions
            frames = torch.randint(0, 256, (4, 256, 256), dtype=torch.uint8, generator=g).cuda()
        else:
            frames = torch.from_numpy(_clip(*shape)).cuda()
        # mode 2: 8 stem warps, bf16 im2col operand; mode 3 (default): 16 stem warps, f16 im2col
        # operand built by byte permutes (u8 values are exact in either type), other K order
        for mode in (2, 3):
            try:
                native_model.fuse_stem = 1
                ref = native_model.run(frames, want_logits=True)
                native_model.fuse_stem = mode
                got = native_model.run(frames, want_logits=True)
                again = native_model.run(frames, want_logits=True)
            finally:
                native_model.fuse_stem = default
            assert all(torch.equal(a, b) for a, b in zip(got, again)), (shape, mode)   # deterministic
            d = (ref[0] - got[0]).abs()
            scale = max(1.0, ref[0].abs().max().item())
            print(shape, "tc stem mode", mode, "vs fp32 stem: max|dz|", d.max().item(), "mean",
                  d.mean().item())
            assert d.max().item() <= 2e-2 * scale and d.mean().item() <= 1e-3 * scale, (shape, mode)
            assert (ref[1] != got[1]).sum().item() <= 1e-4 * ref[1].numel() + 2, (shape, mode)
            assert torch.equal(got[2].cpu(),
                               (got[1] > 0).flatten(1).sum(1).to(torch.int32).cpu()), (shape, mode)


def test_repeated_launch_is_idempotent(eager_model):
    native_model = eager_model
    """ogl_unet_set_repeat (the energy-measurement aid): enqueueing one launch of the schedule
    several times leaves logits and masks unchanged -- every launch reads and writes distinct
    tensors -- for each kind of launch (fused stem, conv, transposed conv, composed level-0 conv)."""
    from openglottal_b200 import _native

    lib = _native.load()
    frames = torch.from_numpy(_clip(3)).cuda()
    ref = native_model.run(frames, want_logits=True)
    n_launch = lib.ogl_unet_launch_count(native_model._handle)
    try:
        for idx in (0, 1, 9, 11, n_launch - 2):   # fused stem, conv, composed convT+conv (x2), level 0
            _native.check(lib.ogl_unet_set_repeat(native_model._handle, idx, 3))
            got = native_model.run(frames, want_logits=True)
            assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1]), idx
            assert torch.equal(ref[2], got[2]), idx
        assert lib.ogl_unet_set_repeat(native_model._handle, 0, 0) != 0      # times out of range
        # the head launch adds popcounts to the area vector: repeating it with an area output
        # would multiply the areas, so the forward refuses (and runs without one)
        _native.check(lib.ogl_unet_set_repeat(native_model._handle, n_launch - 1, 2))
        with pytest.raises(RuntimeError, match="head launch is repeated"):
            native_model.run(frames)
        got = native_model.run(frames, want_logits=True, want_area=False)
        assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
    finally:
        _native.check(lib.ogl_unet_set_repeat(native_model._handle, -1, 1))


def test_bf16_matches_reference_within_north_star(native_model, trained_sd):
    """north_star: logits max-abs <= 2e-2 in bf16. bf16 operands carry 8 mantissa bits, and IDEAL
    bf16 arithmetic (the CPU bit-model: fp32 accumulation, every stored activation and weight
    rounded once) already misses 2e-2 at large |z| on these weights (0.025 on this clip, 0.012 for
    |z| < 1), so the overall bound is tied to that floor, measured on the same frames: the kernel
    may not be worse than 1.05 x the bit-model's own distance from the fp32 reference (or 2e-2,
    whichever is larger). Near the decision boundary (|z| < 1) the 2e-2 bar holds as stated; the
    f16-operand mode (next test) meets it everywhere."""
    from oracle import unet_oracle as uo
    from openglottal_b200 import dice

    frames = _clip(16)
    ref_lg, ref_mask, ref_area = uo.batch_masks(trained_sd, frames)
    bit = uo.folded_forward(trained_sd, uo.frames_to_input(frames), bf16=True)[:, 0].numpy()
    lg, mask, area = native_model.run(torch.from_numpy(frames).cuda(), want_logits=True)
    lg, mask, area = lg.cpu().numpy(), mask.cpu().numpy(), area.cpu().numpy()
    err = np.abs(lg - ref_lg)
    floor = np.abs(bit - ref_lg).max()
    near = np.abs(ref_lg) < 1.0
    print(f"bf16 vs fp32 oracle: max|err|={err.max():.4g} (|z|<1: {err[near].max() if near.any() else 0:.4g}) "
          f"mean={err.mean():.3g}; ideal-bf16 bit-model vs fp32 oracle: max={floor:.4g}; "
          f"kernel vs bit-model: max={np.abs(lg - bit).max():.4g}; logit std={ref_lg.std():.3g}")
    assert err[near].max() <= 2e-2 if near.any() else True
    assert err.max() <= max(2e-2, 1.05 * floor)
    d = dice(mask, ref_mask)
    print("dice", d, "areas", area[:6], ref_area[:6])
    assert d >= 0.999
    assert set(np.unique(mask)) <= {0, 255}
    # popcount of the mask the kernel itself produced is bit-exact
    assert np.array_equal(area, (mask > 0).reshape(len(mask), -1).sum(1))
    rel = np.abs(area - ref_area) / np.maximum(ref_area, 1)
    assert rel.max() <= 0.005


def test_fp16_operand_mode_meets_2e2_everywhere(lib, trained_sd):
    """The same kernels compiled for f16 operands (precision "fp16": tcgen05 kind::f16 with f16
    A/B, satfinite conversions): logits within 2e-2 of the fp32 reference OVERALL (north_star's bar
    as stated), tight against the f16 bit-model, and the bf16 mask / area bars."""
    import openglottal_b200 as ogl
    from oracle import unet_oracle as uo

    m = ogl.UNet().to("cuda")
    m.load_state_dict(trained_sd)
    m.eval()
    m.precision = "fp16"
    frames = _clip(16)
    ref_lg, ref_mask, ref_area = uo.batch_masks(trained_sd, frames)
    bit = uo.folded_forward(trained_sd, uo.frames_to_input(frames), bf16=True, operand="fp16")[:, 0].numpy()
    lg, mask, area = m.run(torch.from_numpy(frames).cuda(), want_logits=True)
    lg, mask, area = lg.cpu().numpy(), mask.cpu().numpy(), area.cpu().numpy()
    err = np.abs(lg - ref_lg)
    print(f"f16 operands vs fp32 oracle: max|err|={err.max():.4g} mean={err.mean():.3g}; "
          f"kernel vs f16 bit-model: max={np.abs(lg - bit).max():.4g}")
    assert err.max() <= 2e-2
    assert np.abs(lg - bit).max() <= 1e-2
    assert ogl.dice(mask, ref_mask) >= 0.999
    assert np.array_equal(area, (mask > 0).reshape(len(mask), -1).sum(1))
    assert (np.abs(area - ref_area) / np.maximum(ref_area, 1)).max() <= 0.005
    # every schedule variant of the f16 twins: fp32 input (separate stem), CUDA-core fused stem,
    # direct full-resolution level, forced CTA pairs -- all within the same bar
    dev = torch.from_numpy(frames[:4]).cuda()
    base = m.run(dev, want_logits=True)[0]
    variants = []
    variants.append(m.run(dev.float() / 255.0, want_logits=True)[0])
    for attr, val in (("fuse_stem", 1), ("fuse_stem", 0), ("schedule", "direct"), ("cta_pairs", 3)):
        old = getattr(m, attr)
        setattr(m, attr, val)
        try:
            variants.append(m.run(dev, want_logits=True)[0])
        finally:
            setattr(m, attr, old)
    for v in variants:
        assert (v - base).abs().max().item() <= 1e-2
        assert np.abs(v.cpu().numpy() - ref_lg[:4]).max() <= 2e-2
    # bf16 stays the default of a fresh model and is unaffected by the f16 pack
    m.precision = "bf16"
    lg_b = m.run(dev, want_logits=True)[0]
    assert (lg_b.cpu() - torch.from_numpy(ref_lg[:4])).abs().max().item() <= 5e-2


def test_forward_signature_matches_reference(native_model, trained_sd):
    """model(x) with (N,1,H,W) f32 returns (N,1,H,W) f32 logits like unet.py:74-88."""
    from oracle import unet_oracle as uo

    x = uo.frames_to_input(_clip(2, 64, 96)).cuda()
    with torch.no_grad():
        y = native_model(x)
    assert y.shape == (2, 1, 64, 96) and y.dtype == torch.float32
    ref = uo.ref_forward(trained_sd, x.cpu())
    assert (y.cpu() - ref).abs().max() <= 5e-2


def test_half_precision_input_is_widened(native_model):
    x = torch.from_numpy(_clip(2, 64, 96)).cuda().float() / 255.0
    xb = x.to(torch.bfloat16)
    with torch.no_grad():
        y16 = native_model(xb.unsqueeze(1))
        y32 = native_model(xb.float().unsqueeze(1))
    assert y16.dtype == torch.float32 and torch.equal(y16, y32)
    with pytest.raises(TypeError):
        native_model.run(torch.zeros((1, 32, 32), dtype=torch.int32, device="cuda"))


def test_batch_chunking_and_ragged_tail(native_model):
    """Odd batch sizes (tail tiles) and chunked calls give identical results."""
    frames = torch.from_numpy(_clip(7)).cuda()
    _, m_all, a_all = native_model.run(frames)
    native_model.max_batch = 3
    try:
        _, m_chunk, a_chunk = native_model.run(frames)
    finally:
        native_model.max_batch = 512
    assert torch.equal(m_all, m_chunk) and torch.equal(a_all, a_chunk)
    _, m1, a1 = native_model.run(frames[:1])
    assert torch.equal(m1[0], m_all[0]) and a1[0] == a_all[0]


def test_unet_segment_frame_reference_semantics(native_model, trained_sd):
    """utils.py:218-241 incl. the resize path for non-256 frames (512x256: exact 2:1)."""
    from oracle import unet_oracle as uo
    import openglottal_b200 as ogl

    f256 = _clip(1)[0]
    got = ogl.unet_segment_frame(f256, native_model, torch.device("cuda"))
    ref = uo.segment_frame(trained_sd, f256)
    assert got.shape == ref.shape and got.dtype == np.uint8
    assert ogl.dice(got, ref) >= 0.999
    f512 = _clip(1, 512, 256)[0]
    got = ogl.unet_segment_frame(f512, native_model, torch.device("cuda"))
    ref = uo.segment_frame(trained_sd, f512)
    assert got.shape == (512, 256)
    d512 = ogl.dice(got, ref)
    print('512x256 single frame dice', d512, 'differing pixels', int((got != ref).sum()))
    assert d512 >= 0.999 or (got != ref).sum() <= 4    # one frame: 4 boundary pixels of ~1600


def test_cuda_graph_replay_equals_eager(native_model):
    """Small batches replay the forward's launches as one CUDA graph from the second call with the
    same signature on: same logits, masks and areas as the eager launches, bit for bit; outputs are
    fresh tensors (a later call does not overwrite an earlier result)."""
    frames = torch.from_numpy(_clip(6)).cuda()
    other = torch.from_numpy(_clip(6, seed=9)).cuda()
    native_model.use_graphs = False
    try:
        ref = native_model.run(frames, want_logits=True)
        ref_other = native_model.run(other, want_logits=True)
    finally:
        native_model.use_graphs = True
    native_model._graphs = {}
    first = native_model.run(frames, want_logits=True)        # eager (first sight of the signature)
    second = native_model.run(frames, want_logits=True)       # captured, then replayed
    third = native_model.run(other, want_logits=True)         # replayed with other frames
    key = next(iter(native_model._graphs))
    assert native_model._graphs[key][1] is not None
    for got in (first, second):
        assert all(torch.equal(a, b) for a, b in zip(ref, got))
    assert all(torch.equal(a, b) for a, b in zip(ref_other, third))
    assert all(torch.equal(a, b) for a, b in zip(ref, second))    # not overwritten by the third call


def test_errors_are_loud(native_model):
    import openglottal_b200 as ogl

    with pytest.raises(ValueError):
        native_model.run(torch.zeros((1, 100, 256), dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError):
        native_model.run(torch.zeros((1, 256, 256), dtype=torch.uint8))  # CPU tensor
    native_model.train()
    try:
        with pytest.raises(RuntimeError):
            native_model.run(torch.zeros((1, 256, 256), dtype=torch.uint8, device="cuda"))
    finally:
        native_model.eval()
    with pytest.raises(NotImplementedError):
        ogl.UNet(1, 1, (16, 32))
    cpu_model = ogl.UNet()
    with pytest.raises(RuntimeError):
        cpu_model.eval()(torch.zeros(1, 1, 32, 32))


def test_segment_clip_host_paths_agree(native_model):
    """Pinned input (copied from in place), pageable input (staged through two pinned buffers) and
    device input give identical areas and masks; ragged last batch included."""
    import openglottal_b200 as ogl

    frames = torch.from_numpy(_clip(11))
    a_dev, m_dev = ogl.segment_clip(frames.cuda(), native_model, batch=4, want_masks=True)
    a_page, m_page = ogl.segment_clip(frames, native_model, batch=4, want_masks=True)
    a_pin, m_pin = ogl.segment_clip(frames.pin_memory(), native_model, batch=4, want_masks=True)
    assert torch.equal(a_dev, a_page) and torch.equal(a_dev, a_pin)
    assert torch.equal(m_dev, m_page) and torch.equal(m_dev, m_pin)
