"""Error behaviour and cache semantics of the drop-in boundary (through the C ABI, on a B200):
the reference's NaN guard, argument validation before any raw pointer reaches a kernel, weight staleness, engine LRU.
"""
import pytest
import torch

from helpers import D, gc

pytestmark = pytest.mark.gpu
H, W = 128, 192


def _inputs(B=1, seed=5):
    frames, masks = D.clips.synthetic_clip(seed, B, 3, H, W)
    return frames.cuda(), masks.cuda()


def _model(variant="fast"):
    torch.manual_seed(gc.SEED_P)
    return D.build_p_model(variant).eval().cuda()


def test_nan_guard_raises_like_the_reference():
    """seg_video_model_fast.py:152-156: `[NaNGuard] non-finite activations after <tag>` as a RuntimeError.  Here the
    check is one fused launch per frame; the error surfaces at check_finite() / a later call (default) or at the call
    itself (strict_finite)."""
    fr, mk = _inputs()
    m = _model("fast")
    x = torch.cat([fr[:, 1], mk[:, 1]], 1)
    dpb = {"frame": fr[:, 0], "feature": None}
    with torch.no_grad():
        m(x, 32, dpb, after_i=True)
        m.check_finite()                                    # a healthy frame raises nothing
        m.feature_extractor.conv2[3].ffn[2].bias.data[5] = float("nan")
        m.invalidate_weights()
        m(x, 32, dpb, after_i=True)
        with pytest.raises(RuntimeError, match=r"\[NaNGuard\] non-finite activations after .*feature_extractor.ctx"):
            m.check_finite()
        m.strict_finite = True
        with pytest.raises(D.NonFiniteError):
            m(x, 32, dpb, after_i=True)


def test_fp16_range_limit_is_flagged():
    """The split storage format ends at 65504: a saturated activation counts as non-finite instead of passing silently."""
    fr, mk = _inputs()
    m = _model("old")
    with torch.no_grad():
        m.feature_adaptor_i.adaptor.weight.mul_(1e6)
        m(fr[:, 1], 32, {"frame": fr[:, 0], "feature": None}, after_i=True)
        with pytest.raises(D.NonFiniteError, match="feature_adaptor"):
            m.check_finite()


def test_dpb_and_input_validation():
    fr, mk = _inputs()
    m = _model("performance")
    x = torch.cat([fr[:, 1], mk[:, 1]], 1)
    good = {"frame": fr[:, 0], "feature": None}
    with torch.no_grad():
        r = m(x, 32, good, after_i=True)
        with pytest.raises(RuntimeError, match="shape"):
            m(x, 32, {"frame": fr[:, 0, :, :64], "feature": None}, after_i=True)
        with pytest.raises(TypeError, match="float32"):
            m(x, 32, {"frame": fr[:, 0].half(), "feature": None}, after_i=True)
        with pytest.raises(RuntimeError, match="is on cpu"):
            m(x, 32, {"frame": fr[:, 0].cpu(), "feature": None}, after_i=True)
        with pytest.raises(RuntimeError, match="shape"):
            m(x, 32, {"frame": None, "feature": r["dpb"]["feature"][:, :128]}, after_i=False)
        with pytest.raises(RuntimeError, match="None"):
            m(x, 32, {"frame": fr[:, 0], "feature": None}, after_i=False)
        with pytest.raises(RuntimeError, match="channels"):
            m(torch.cat([x, x[:, :1]], 1), 32, good, after_i=True)
        with pytest.raises(RuntimeError, match="multiples of 64"):
            m(x[:, :, :80], 32, {"frame": fr[:, 0, :, :80], "feature": None}, after_i=True)
        with pytest.raises(NotImplementedError):
            m.train()(x, 32, good, after_i=True)
        m.eval()
        r2 = m(x, 32, good, after_i=True)                  # the module still works after every refusal
    assert torch.equal(r["dpb"]["frame"], r2["dpb"]["frame"])


def test_weight_edits_reach_the_engine():
    """In-place edits (version counter), re-registered Parameters (registration hook), `.data` writes (device
    checksum / invalidate_weights) and load_state_dict all repack before the next forward uses the weights."""
    fr, mk = _inputs()
    x, dpb = fr[:, 1], {"frame": fr[:, 0], "feature": None}

    def run(m):
        with torch.no_grad():
            return m(x, 32, dpb, after_i=True)["dpb"]["frame"].clone()

    m = _model("old")
    base = run(m)
    ref = _model("old")
    target = ref.recon_generation_net.head.weight
    with torch.no_grad():
        target.mul_(0.5)
    want = run(ref)
    assert not torch.equal(base, want)

    with torch.no_grad():                                   # (1) in-place op: version counter
        m.recon_generation_net.head.weight.mul_(0.5)
    assert torch.equal(run(m), want)
    m.recon_generation_net.head.weight = torch.nn.Parameter(m.recon_generation_net.head.weight.detach() * 2.0)
    assert torch.equal(run(m), base)                        # (2) re-registered Parameter: seen at once
    m.recon_generation_net.head.weight.data.copy_(target.data)      # (3) write through .data: invisible to (1), (2)
    m.checksum_every = 1
    assert torch.equal(run(m), want)
    m.checksum_every = 0
    m.recon_generation_net.head.weight.data.mul_(2.0)
    assert torch.equal(run(m), want)                        # the documented blind spot without the checksum ...
    m.invalidate_weights()
    assert torch.equal(run(m), base)                        # ... closed by invalidate_weights()
    m.load_state_dict(ref.state_dict())                     # (4) load_state_dict
    assert torch.equal(run(m), want)


def test_engine_cache_is_bounded():
    m = _model("old")
    m.max_engines = 2
    with torch.no_grad():
        for h, w in ((64, 64), (64, 128), (128, 128), (64, 64)):
            f = torch.rand(1, 3, h, w, device="cuda")
            r = m(f, 32, {"frame": f, "feature": None}, after_i=True)
            assert r["dpb"]["frame"].shape == (1, 3, h, w)
            assert len(m._engines) <= 2
    m.release_engines()
    assert len(m._engines) == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_non_current_device():
    """model.to('cuda:1') without torch.cuda.set_device(1): every C-ABI call runs on the engine's device."""
    fr, mk = _inputs()
    m0 = _model("old")
    with torch.no_grad():
        want = m0(fr[:, 1], 32, {"frame": fr[:, 0], "feature": None}, after_i=True)["dpb"]["frame"].cpu()
        m1 = _model("old").to("cuda:1")
        assert torch.cuda.current_device() == 0
        got = m1(fr[:, 1].to("cuda:1"), 32, {"frame": fr[:, 0].to("cuda:1"), "feature": None}, after_i=True)
    assert torch.equal(got["dpb"]["frame"].cpu(), want)


def test_graph_replay_matches_direct_launches(monkeypatch):
    """A forward is captured once per (after_i, qp, mask present) as a CUDA graph and replayed on whatever tensors the
    next call brings (caller pointers live in device slots).  Replays on fresh tensors give the bits of direct launches."""
    fr, mk = _inputs(seed=9)

    def gop(m):
        outs = []
        with torch.no_grad():
            dpb = {"frame": fr[:, 0].clone(), "feature": None}
            for rep in range(2):                    # second pass: every graph is a replay, every tensor a new one
                d = dpb
                for t, qp in ((1, 40), (2, 32), (2, 36), (1, 32)):
                    x = torch.cat([fr[:, t], mk[:, t]], 1).clone() if t == 2 else fr[:, t].clone()   # with / without mask
                    r = m(x, qp, d, after_i=(t == 1))
                    outs.append((r["dpb"]["frame"].clone(), r["dpb"]["feature"].clone(), r["bpp"].clone()))
                    if t == 1:
                        d = r["dpb"]
        m.check_finite()
        return outs

    for variant in ("performance", "mask_prop"):
        a = gop(_model(variant))
        monkeypatch.setenv("DMC_GRAPH", "0")
        b = gop(_model(variant))
        monkeypatch.delenv("DMC_GRAPH")
        for (xa, fa, ba), (xb, fb, bb) in zip(a, b):
            assert torch.equal(xa, xb) and torch.equal(fa, fb)
            assert float((ba - bb).abs().max()) <= 1e-6 * float(bb.abs().max())
        n = len(a) // 2
        for i in range(n):                          # pass 2 == pass 1
            assert torch.equal(a[i][0], a[n + i][0])


def test_two_streams_do_not_trap_the_chain_kernel():
    """The persistent chain kernel waits on tiles owned by other clusters of its grid, so it needs the whole grid
    resident; it is launched cooperatively, which makes the driver start a grid only when all of it fits.  Two modules
    driven from two streams at the same time (the scenario that could otherwise leave two half-resident grids waiting
    for each other until the bounded waits trap) finish, with the results of a sequential run."""
    fr, mk = _inputs(seed=21)
    x = torch.cat([fr[:, 1], mk[:, 1]], 1)
    dpb = {"frame": fr[:, 0], "feature": None}
    ma, mb = _model("performance"), _model("fast")
    with torch.no_grad():
        want_a = ma(x, 32, dpb, after_i=True)["dpb"]["frame"].clone()
        want_b = mb(x, 40, dpb, after_i=True)["dpb"]["frame"].clone()
        torch.cuda.synchronize()
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        outs_a, outs_b = [], []
        for _ in range(12):
            with torch.cuda.stream(sa):
                outs_a.append(ma(x, 32, dpb, after_i=True)["dpb"]["frame"])
            with torch.cuda.stream(sb):
                outs_b.append(mb(x, 40, dpb, after_i=True)["dpb"]["frame"])
        torch.cuda.synchronize()
    assert all(torch.equal(o, want_a) for o in outs_a) and all(torch.equal(o, want_b) for o in outs_b)
