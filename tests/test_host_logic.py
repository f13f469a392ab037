"""Host-side logic that needs no GPU: the drop-in modules' parameter trees and mode checks, the
synthetic clip generator, clip sharding and the statistics formulas."""
import math

import pytest
import torch

from helpers import D, O, gc

clips = D.clips

EXPECTED = {"old": (341, 20691456), "performance": (378, 22919168), "fast": (347, 20712608),
            "mask_prop": (355, 21194593), "intra": (407, 45651314)}


@pytest.mark.parametrize("variant", list(EXPECTED))
def test_state_dict_layout(variant):
    m = D.DMCI() if variant == "intra" else D.build_p_model(variant)
    sd = m.state_dict()
    assert (len(sd), sum(v.numel() for v in sd.values())) == EXPECTED[variant]
    assert not list(m.buffers())
    keys = set(sd)
    if variant == "old":
        assert "encoder.conv3.dc.0.weight" in keys and "decoder.conv2.weight" in keys
    elif variant != "intra":
        assert "encoder.conv2.2.dc.0.weight" in keys and "decoder.proj.weight" in keys
        assert "hyper_in_adapter.weight" in keys
    if variant == "performance":
        assert sd["q_sft"].shape == (72, 256, 1, 1) and "mask_sft.down.weight" in keys
    if variant == "mask_prop":
        assert sd["mask_predictor.net.0.weight"].shape == (64, 768, 3, 3)
    if variant == "intra":
        assert sd["y_prior_fusion.3.weight"].shape == (514, 512, 1, 1)
    # optimizer grouping of the trainer keys off these substrings (trainer:573-591)
    names = [n for n, _ in m.named_parameters()]
    assert any("bit_estimator" in n for n in names)


def test_identity_adapter_init():
    w = D.DMC_fast().hyper_in_adapter.weight
    assert torch.equal(w[:, :128, 0, 0], torch.eye(128)) and float(w[:, 128].abs().sum()) == 0.0


def test_shift_qp_and_modes():
    m = D.DMC_old()
    assert [m.shift_qp(32, i) for i in range(3)] == [32, 40, 36]
    x = torch.rand(1, 3, 64, 64)
    with pytest.raises(RuntimeError):
        m(x, 32, {"frame": x, "feature": None})          # CPU tensor
    assert D.build_p_model("fast").variant == "fast"
    with pytest.raises(ValueError):
        D.build_p_model("nope")


def test_synthetic_clip_properties():
    f, m = clips.synthetic_clip(5, 2, 4, 128, 192)
    f2, m2 = clips.synthetic_clip(5, 2, 4, 128, 192)
    assert torch.equal(f, f2) and torch.equal(m, m2)
    assert f.shape == (2, 4, 3, 128, 192) and m.shape == (2, 4, 1, 128, 192)
    assert 0.0 <= float(f.min()) and float(f.max()) <= 1.0
    assert set(m.unique().tolist()) <= {0.0, 1.0}
    assert 0.02 < float(m.mean()) < 0.5
    # consecutive frames are shifted copies: far more alike than independent noise
    assert float((f[:, 1] - f[:, 0]).abs().mean()) < float((f[:, 1] - f[:, 0].flip(-1)).abs().mean())
    assert not torch.equal(clips.synthetic_clip(6, 1, 2, 64, 64)[0], clips.synthetic_clip(5, 1, 2, 64, 64)[0])


def test_shard_clips_partitions():
    for world in (1, 2, 4, 8):
        parts = [clips.shard_clips(64, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(map(len, parts)) - min(map(len, parts)) == 0
    assert clips.shard_clips(5, 1, 2) == [1, 3]
    assert clips.shard_clips(0, 0, 2) == []


def test_clip_stats_cpu_formulas_match_trainer_metrics():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 64, 64, generator=g)
    xh = (x + 0.05 * torch.randn(2, 3, 64, 64, generator=g)).clamp(0, 1)
    mask = torch.zeros(2, 1, 64, 64)
    mask[:, :, 10:30, 5:40] = 1
    res = {"dpb": {"frame": xh}, "bpp": torch.tensor([1.5, 2.5]), "bpp_y": torch.tensor([1.0, 2.0]),
           "bpp_z": torch.tensor([0.5, 0.5])}
    st = clips.ClipStats("cpu")
    st.add_frame(res, x, mask)
    s = st.summary()
    assert s["frames"] == 2
    assert abs(s["bpp"] - 2.0) < 1e-9 and abs(s["bpp_z"] - 0.5) < 1e-9
    assert abs(s["psnr"] - float(O.psnr_from_mse(O.mse(xh, x)))) < 1e-4
    assert abs(s["roi_psnr"] - float(O.psnr_from_mse(O.roi_mse(xh, x, mask)))) < 1e-4
    empty = clips.ClipStats("cpu")
    empty.add_frame(res, x, torch.zeros_like(mask))
    assert abs(empty.summary()["roi_psnr"] - empty.summary()["psnr"]) < 1e-12   # trainer:656-657 fallback


def test_gop_qp_schedule():
    assert [clips.gop_qp(32, t) for t in range(1, 9)] == [40, 32, 36, 32, 36, 32, 36, 32]
    assert [O.shift_qp(32, O.INDEX_MAP[t % 8]) for t in range(1, 9)] == [clips.gop_qp(32, t) for t in range(1, 9)]


def test_weight_signature_cache_sees_edits_moves_and_reloads():
    """The drop-in modules re-pack their weights when a parameter changes: the signature is built from a cached
    parameter list (walking the module tree costs ~1 ms per forward) and must still see in-place edits, `.to()` /
    `load_state_dict`, and -- immediately, through the global registration hook -- a re-registered Parameter."""
    m = D.build_p_model("old").eval()
    s0 = m._signature()
    assert m._signature() == s0
    with torch.no_grad():
        next(m.parameters()).add_(1.0)                      # in-place edit: version counter
    s1 = m._signature()
    assert s1 != s0
    m.load_state_dict(m.state_dict())                       # copies in place and drops the cache
    s2 = m._signature()
    assert s2 != s1
    m = m.double().float()                                  # _apply: new storages
    assert m._signature() != s2
    name, mod = next((n, sub) for n, sub in m.named_modules() if getattr(sub, "weight", None) is not None)
    s3 = m._signature()
    mod.weight = torch.nn.Parameter(mod.weight.detach().clone())   # re-registered: the very next signature differs
    assert m._signature() != s3


def test_writes_through_data_need_the_checksum_or_invalidate():
    """`p.data.copy_()` (trainer_seg_video_model.py:789-791) bumps no version counter and keeps the pointer: the
    signature cannot see it.  The periodic device checksum does (GPU test), and `invalidate_weights()` forces the
    repack at once."""
    m = D.build_p_model("old").eval()
    s0 = m._signature()
    p = next(m.parameters())
    p.data.copy_(p.data + 1.0)
    assert m._signature() == s0                             # the blind spot the checksum exists for
    m._weights_sig[("key",)] = s0
    m.invalidate_weights()
    assert m._weights_sig == {}


def test_modules_copy_and_pickle_without_engine_state():
    """Engine handles / the ctypes library are process-local: deepcopy and pickle drop them (a copied raw handle
    would be destroyed twice)."""
    import copy
    import pickle
    m = D.build_p_model("fast").eval()
    m._engines[(1, 64, 64, 0, 0)] = 12345                   # a fake handle: must not travel
    m.__dict__["_lib"] = None
    c = copy.deepcopy(m)
    assert c._engines == {} and c is not m
    assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), m.state_dict().values()))
    r = pickle.loads(pickle.dumps(m))
    assert r._engines == {} and set(r.state_dict()) == set(m.state_dict())
    m._engines.clear()


def test_reference_copy_matches_shim_initialisation():
    """oracle/_ref (the unmodified reference modules, copied by oracle/make_ref.py) and the drop-in modules give the
    same parameters under the same seed: same keys, shapes and values."""
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    R = make_ref.load_reference()
    for variant in ("old", "performance", "fast", "mask_prop"):
        torch.manual_seed(7)
        ref = R[variant]().state_dict()
        torch.manual_seed(7)
        mine = D.build_p_model(variant).state_dict()
        assert list(ref) == list(mine)
        assert all(torch.equal(ref[k], mine[k]) for k in ref)
