"""Pins oracle/dmc_oracle.py to the real reference: the fixtures in tests/golden were produced by
oracle/make_golden.py running the unmodified reference modules from /root/reference."""
import numpy as np
import pytest
import torch

from helpers import O, gc, golden, sd_of, seeded_models


@pytest.mark.parametrize("case", gc.CASES, ids=lambda c: c["name"])
def test_default_init_matches_reference_checksums(case):
    g = golden(case["name"])
    mi, _ = seeded_models("old", case)
    np.testing.assert_allclose(gc.sd_checksum(mi.state_dict()), g["sd_checksum_intra"], rtol=0, atol=0)
    for variant in gc.case_variants(case):
        _, mp = seeded_models(variant, case)
        np.testing.assert_allclose(gc.sd_checksum(mp.state_dict()), g[f"sd_checksum_{variant}"], rtol=0, atol=0)


def _check(g, tag, res, taps, target, mask):
    bpp3 = torch.stack([res["bpp"], res["bpp_y"], res["bpp_z"]], 1).numpy()
    np.testing.assert_allclose(bpp3, g[f"{tag}/bpp3"], rtol=1e-6, atol=0)
    assert np.array_equal(taps["y_q"].numpy().astype(np.int8), g[f"{tag}/y_q"])
    assert np.array_equal(taps["z_hat"].numpy().astype(np.int8), g[f"{tag}/z_hat"])
    x_hat = res["dpb"]["frame"]
    np.testing.assert_allclose(x_hat[:, :, ::8, ::8].numpy(), g[f"{tag}/x_hat_sub"], rtol=0, atol=1e-6)
    p, r = gc.metrics(x_hat, target, mask)
    np.testing.assert_allclose([p, r], g[f"{tag}/psnr"], rtol=0, atol=1e-5)
    if res["dpb"].get("feature") is not None:
        np.testing.assert_allclose(res["dpb"]["feature"][:, :, ::4, ::4].numpy(), g[f"{tag}/feature_sub"],
                                   rtol=0, atol=1e-5)
    np.testing.assert_allclose(taps["scales_hat"][:, :, ::2, ::2].numpy(), g[f"{tag}/scales_sub"], rtol=0, atol=1e-5)
    if res.get("mask_pred") is not None:
        np.testing.assert_allclose(res["mask_pred"][:, :, ::8, ::8].numpy(), g[f"{tag}/mask_pred_sub"],
                                   rtol=0, atol=1e-5)


@pytest.mark.parametrize("case", gc.CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("variant", gc.VARIANTS)
def test_oracle_reproduces_reference(case, variant):
    if variant not in gc.case_variants(case):
        pytest.skip("the reference does not pad y in this variant: it cannot run this size")
    g = golden(case["name"])
    frames, masks = gc.case_inputs(case)
    mi, mp = seeded_models(variant, case)
    sd_i, sd_p = sd_of(mi), sd_of(mp)
    taps = {}
    r_i = O.dmci_forward(sd_i, frames[:, 0], case["qp"], taps)
    if variant == "old":        # the intra frame is shared by all variants; check it once
        _check(g, "intra/0", r_i, taps, frames[:, 0], None)
    dpb = r_i["dpb"]
    for t in range(1, frames.shape[1]):
        qp = O.shift_qp(case["qp"], O.INDEX_MAP[t % 8])
        x_in = frames[:, t] if variant == "old" else torch.cat([frames[:, t], masks[:, t]], 1)
        taps = {}
        r = O.dmc_forward(sd_p, variant, x_in, qp, dpb, after_i=(t == 1), taps=taps)
        dpb = r["dpb"]
        _check(g, f"{variant}/{t}", r, taps, frames[:, t], masks[:, t])


def test_anchor_table_of_survey():
    """SURVEY.md section 4 / BASELINE.md known-answer bpp values."""
    g = golden("anchor_256")
    expect = {"old": (2.427825, 1.903015, 1.845709), "performance": (5.096867, 4.769346, 4.666419),
              "fast": (3.958817, 3.177323, 3.065238), "mask_prop": (3.958817, 3.222866, 3.047088)}
    assert abs(float(g["intra/0/bpp3"][0, 0]) - 3.932585) < 2e-6
    for v, vals in expect.items():
        for t, e in enumerate(vals, start=1):
            assert abs(float(g[f"{v}/{t}/bpp3"][0, 0]) - e) < 2e-6
