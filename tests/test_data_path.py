"""Device data path: the reference's per-frame preparation -- uint8 camera frame -> BT.709
YCbCr in [0,1], cached mask as channel 4, one crop per sequence (src/dataset/seg_waymo_dataset.py:26-43,56-79,231-245).

CPU:  oracle/data_oracle.py against the fixture minted from the reference's own functions (bit for bit).
GPU:  csrc k_frames_from_u8 through the C ABI against the oracle, BIT-EXACT (integer / byte work in, IEEE fp32 in the
      reference's operation order out), plus the mask hand-over of a propagated mask_prop GOP."""
import numpy as np
import pytest
import torch

from helpers import D, golden
from oracle import data_oracle as DO


def test_oracle_matches_the_reference_fixture():
    g = golden("data_path")
    img, want = g["img"], g["ycbcr"]
    for t in range(img.shape[0]):
        got = DO.rgb_to_ycbcr_bt709(DO.rgb_from_u8(img[t])).numpy()
        assert np.array_equal(got, want[t])                        # bit for bit
    full = DO.item(img, g["mask"])
    assert full.shape == (img.shape[0], 4, img.shape[1], img.shape[2])
    assert np.array_equal(full[:, :3].numpy(), want) and np.array_equal(full[:, 3].numpy(), g["mask"].astype(np.float32))
    # the cube corners hit both ends of the clamp and the chroma extremes
    assert float(full[:, :3].min()) == 0.0 and float(full[:, :3].max()) == 1.0


def test_host_side_argument_checks():
    with pytest.raises(RuntimeError):
        D.data.frames_from_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))          # no CPU path
    with pytest.raises(TypeError):
        D.data.mask_from_logits(torch.zeros(4))


@pytest.mark.gpu
def test_fixture_frames_bit_exact():
    g = golden("data_path")
    img, mask = torch.from_numpy(g["img"]).cuda(), torch.from_numpy(g["mask"]).cuda()
    out = D.data.frames_from_u8(img, mask)
    assert torch.equal(out[:, :3].cpu(), torch.from_numpy(g["ycbcr"]))
    assert torch.equal(out[:, 3].cpu(), torch.from_numpy(g["mask"]).float())


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    dict(T=1, H=1280, W=1920, crop=None, bgr=False, thr=0),                  # a whole Waymo FRONT frame
    dict(T=3, H=1280, W=1920, crop=(517, 901, 256, 256), bgr=True, thr=127),   # the trainer's crop, png mask cache, cv2 order
    dict(T=2, H=37, W=53, crop=(3, 5, 30, 41), bgr=False, thr=0),            # ragged: width not a multiple of 4
    dict(T=1, H=16, W=16, crop=(15, 15, 1, 1), bgr=False, thr=0),            # a single pixel at the far corner
], ids=["full_frame", "crop256_bgr_png", "ragged", "one_pixel"])
def test_frames_match_the_oracle_bit_for_bit(case):
    g = np.random.default_rng(case["H"] * 7 + case["W"])
    img = g.integers(0, 256, size=(case["T"], case["H"], case["W"], 3), dtype=np.uint8)
    shape = (case["T"], case["H"], case["W"])
    mask = g.integers(0, 256, size=shape, dtype=np.uint8) if case["thr"] else (g.random(shape) < 0.2).astype(np.uint8)
    want = DO.item(img, mask, crop=case["crop"], bgr=case["bgr"], threshold=case["thr"])
    got = D.data.frames_from_u8(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda(), crop=case["crop"],
                                bgr=case["bgr"], mask_threshold=case["thr"])
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    # without a mask cache the dataset falls back to zeros (strict_masks=False); three-channel output for `old`
    got0 = D.data.frames_from_u8(torch.from_numpy(img).cuda(), None, crop=case["crop"], bgr=case["bgr"])
    assert torch.equal(got0[:, :3].cpu(), want[:, :3]) and float(got0[:, 3].abs().sum()) == 0.0
    got3 = D.data.frames_from_u8(torch.from_numpy(img).cuda(), None, crop=case["crop"], bgr=case["bgr"], with_mask=False)
    assert got3.shape[1] == 3 and torch.equal(got3.cpu(), want[:, :3])


@pytest.mark.gpu
def test_bad_arguments_raise_and_leave_the_library_usable():
    img = torch.zeros(1, 32, 32, 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        D.data.frames_from_u8(img, crop=(0, 0, 64, 64))                      # seg_waymo_dataset.py:235-236
    with pytest.raises(ValueError):
        D.data.frames_from_u8(img, torch.zeros(1, 16, 16, dtype=torch.uint8, device="cuda"))   # :68-69
    with pytest.raises(TypeError):
        D.data.frames_from_u8(img.float())
    assert D.data.frames_from_u8(img).shape == (1, 4, 32, 32)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 5, 4096, 1280 * 1920])
def test_mask_from_logits(n):
    x = torch.randn(n, device="cuda")
    if n > 4:
        x[:4] = torch.tensor([0.0, -0.0, float("nan"), 1e-30], device="cuda")
    assert torch.equal(D.data.mask_from_logits(x), (x > 0).float())


@pytest.mark.gpu
def test_gop_coder_propagates_its_own_masks():
    """clips.run_gop(mask_feedback=True) -- the config-4 protocol -- is what clips.GopCoder packages: after the first
    two P frames the only mask the codec sees is the one its own MaskPredictor produced."""
    H, W, qp = 128, 192, 32
    frames, masks = D.clips.synthetic_clip(9, 1, 5, H, W)
    torch.manual_seed(0)
    mi = D.DMCI().eval().cuda()
    torch.manual_seed(1)
    mp = D.build_p_model("mask_prop").eval().cuda()
    fr, mk = frames.cuda(), masks.cuda()
    ref = D.clips.run_gop(mi, mp, "mask_prop", fr, mk, qp, mask_feedback=True)
    coder = D.clips.GopCoder(mi, mp, qp)
    outs = coder.code(fr, mk[:, 1])
    assert len(outs) == len(ref) == 5
    for a, b in zip(outs, ref):
        assert torch.equal(a["dpb"]["frame"], b["dpb"]["frame"]) and torch.equal(a["bpp"], b["bpp"])
    # masks the coder used: GT for t = 1, 2; its own thresholded prediction from t = 3 on
    assert torch.equal(coder.masks_used[0], mk[:, 1]) and torch.equal(coder.masks_used[1], mk[:, 1])
    assert torch.equal(coder.masks_used[2], (ref[2]["mask_pred"] > 0).float())
    assert set(coder.masks_used[3].unique().tolist()) <= {0.0, 1.0}
