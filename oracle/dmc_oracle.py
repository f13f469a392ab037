"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

CPU restatement (plain torch fp32, functional, state_dict driven) of the
reference's DMC P-frame forward ('old', 'performance', 'fast', 'mask_prop')
and the DMCI intra forward.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this file; the
product (the CUDA engine behind include/dmc_b200.h) never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function
here against fixtures in tests/golden/ that were produced by importing the
*real* reference from /root/reference (generator: oracle/make_golden.py).

Every function cites the reference file:line it restates (paths relative to
the reference root).  The arithmetic order of each elementwise expression is
kept identical to the reference so results are bit-identical on the same
torch build; convolutions go through F.conv2d exactly like nn.Conv2d.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

QP_SHIFT = (0, 8, 4)            # src/refactor/config.py:25, src/models/video_model.py:13
INDEX_MAP = (0, 1, 0, 2, 0, 2, 0, 2)   # trainer_seg_video_model.py:76
VARIANTS = ("old", "performance", "fast", "mask_prop")


# --------------------------------------------------------------------------
# layers  (src/layers/layers.py, src/layers/inference.py)
# --------------------------------------------------------------------------
def wsilu(x: Tensor) -> Tensor:
    """layers.py:8-10."""
    return F.silu(4.0 * x) / 4.0


def wsilu_chunk_add(x: Tensor) -> Tensor:
    """layers.py:12-20."""
    x = wsilu(x)
    a, b = torch.chunk(x, 2, dim=1)
    return a + b


def _conv(sd: SD, key: str, x: Tensor, stride=1, padding=0, groups=1) -> Tensor:
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride,
                    padding=padding, groups=groups)


def depth_conv_block(sd: SD, p: str, x: Tensor, shortcut=False, quant_step=None) -> Tensor:
    """layers.py:43-79 (adaptor present iff its weight is in the state_dict)."""
    if (p + ".adaptor.weight") in sd:
        x = _conv(sd, p + ".adaptor", x)
    t = wsilu(_conv(sd, p + ".dc.0", x))
    t = _conv(sd, p + ".dc.2", t, padding=1, groups=t.shape[1])
    out = _conv(sd, p + ".dc.3", t) + x
    out = _conv(sd, p + ".ffn.2", wsilu_chunk_add(_conv(sd, p + ".ffn.0", out))) + out
    if shortcut:
        out = out + x
    if quant_step is not None:
        out = out * quant_step
    return out


def res_block_stride2(sd: SD, p: str, x: Tensor) -> Tensor:
    """layers.py:81-90."""
    return depth_conv_block(sd, p + ".conv", _conv(sd, p + ".down", x, stride=2), shortcut=True)


def res_block_upsample(sd: SD, p: str, x: Tensor) -> Tensor:
    """layers.py:93-102 with SubpelConv2x (layers.py:22-40), 1x1 kernel."""
    up = F.pixel_shuffle(_conv(sd, p + ".up.conv.0", x), 2)
    return depth_conv_block(sd, p + ".conv", up, shortcut=True)


def _dcb_chain(sd: SD, p: str, n: int, x: Tensor, start: int = 0) -> Tensor:
    for i in range(start, start + n):
        x = depth_conv_block(sd, f"{p}.{i}", x)
    return x


# --------------------------------------------------------------------------
# entropy model pieces
# --------------------------------------------------------------------------
def probs_to_bits(probs: Tensor) -> Tensor:
    """models/common_model.py:30-34."""
    bits = torch.log(probs + 1e-5) * (-1.0 / math.log(2.0))
    return torch.clamp(bits, 0, None)


def gaussian_bits_old(y: Tensor, sigma: Tensor) -> Tensor:
    """models/common_model.py:36-42 (torch Normal.cdf spelled out)."""
    sigma = sigma.clamp(1e-5, 1e10)
    inv = sigma.reciprocal()

    def cdf(v):
        return 0.5 * (1 + torch.erf(v * inv / math.sqrt(2)))

    return probs_to_bits(cdf(y + 0.5) - cdf(y - 0.5))


def gaussian_bits_refactor(y: Tensor, sigma: Tensor) -> Tensor:
    """refactor/common_model.py:37-68."""
    y = torch.nan_to_num(y, nan=0.0, posinf=1e4, neginf=-1e4)
    sigma = torch.nan_to_num(sigma, nan=1e-5, posinf=1e10, neginf=1e-5).clamp(1e-5, 1e10)
    inv = 1.0 / sigma
    z_hi = ((y + 0.5) * inv).clamp(-12.0, 12.0)
    z_lo = ((y - 0.5) * inv).clamp(-12.0, 12.0)
    root2 = math.sqrt(2.0)
    probs = 0.5 * (torch.erf(z_hi / root2) - torch.erf(z_lo / root2))
    probs = torch.nan_to_num(probs, nan=0.0, posinf=0.0, neginf=0.0).clamp_min(1e-9)
    return -torch.log2(probs)


def bitparm_cdf(sd: SD, p: str, v: Tensor, qp: int) -> Tensor:
    """entropy_models.py:84-106 (Bitparm) and :139-150 (get_cdf)."""
    for k in (1, 2, 3, 4):
        h = sd[f"{p}.f{k}.h"][qp:qp + 1]
        b = sd[f"{p}.f{k}.b"][qp:qp + 1]
        v = v * F.softplus(h) + b
        if k < 4:
            a = sd[f"{p}.f{k}.a"][qp:qp + 1]
            v = v + torch.tanh(v) * torch.tanh(a)
    return torch.sigmoid(v)


def z_bits(sd: SD, z_hat: Tensor, qp: int, p: str = "bit_estimator_z") -> Tensor:
    """models/common_model.py:44-47 / refactor/common_model.py:70-73."""
    return probs_to_bits(bitparm_cdf(sd, p, z_hat + 0.5, qp) - bitparm_cdf(sd, p, z_hat - 0.5, qp))


def pad_for_y(y: Tensor) -> Tensor:
    """models/common_model.py:68-72 + inference.py:40-43 (replicate pad to x4)."""
    h, w = y.shape[-2:]
    pb, pr = (-h) % 4, (-w) % 4
    if pb == 0 and pr == 0:
        return y
    return F.pad(y, (0, pr, 0, pb), mode="replicate")


def checkerboard_2x(c: int, h: int, w: int, dtype=torch.float32):
    """models/common_model.py:93-114: mask_0 = ((1,0),(0,1)) on the first
    channel half and its complement on the second half; mask_1 = 1 - mask_0."""
    hh = torch.arange(h).view(1, h, 1)
    ww = torch.arange(w).view(1, 1, w)
    half = (torch.arange(c).view(c, 1, 1) >= c // 2).long()
    m0 = (((hh + ww + half) % 2) == 0).to(dtype).unsqueeze(0)
    return m0, 1.0 - m0


def checkerboard_4x(c: int, h: int, w: int, dtype=torch.float32):
    """models/common_model.py:152-169: the four 2x2 positions rotated over the
    four channel quarters."""
    hh = (torch.arange(h) % 2).view(1, h, 1)
    ww = (torch.arange(w) % 2).view(1, 1, w)
    pos = hh * 2 + ww                                   # 0:(0,0) 1:(0,1) 2:(1,0) 3:(1,1)
    quarter = (torch.arange(c) // (c // 4)).view(c, 1, 1)
    # step s owns position table[s][quarter]
    table = torch.tensor([[0, 1, 2, 3], [3, 2, 1, 0], [2, 3, 0, 1], [1, 0, 3, 2]])
    masks = []
    for s in range(4):
        own = table[s][quarter.view(-1)].view(c, 1, 1)
        masks.append((pos == own).to(dtype).unsqueeze(0))
    return masks


def _process_with_mask(y, scales, means, mask):
    """models/common_model.py:81-90 in eval mode (quant = torch.round)."""
    scales_hat = scales * mask
    means_hat = means * mask
    y_res = (y - means_hat) * mask
    y_q = torch.round(y_res) * mask
    return y_q, y_q + means_hat, scales_hat


def compress_prior_2x(sd: SD, y: Tensor, params: Tensor, taps: Optional[dict] = None):
    """models/common_model.py:121-149 == refactor/common_model.py:147-188 (fm_s=None).
    Returns (symbols, y_hat, scales_hat)."""
    q_dec, scales, means = params.chunk(3, 1)
    q_dec = torch.clamp_min(q_dec, 0.5)                    # inference.py:29-33
    y = y * torch.reciprocal(q_dec)
    m0, m1 = checkerboard_2x(y.shape[1], y.shape[2], y.shape[3], y.dtype)
    q0, yh0, s0 = _process_with_mask(y, scales, means, m0)
    sp = spatial_prior(sd, torch.cat((yh0, params), dim=1))
    scales1, means1 = sp.chunk(2, 1)
    q1, yh1, s1 = _process_with_mask(y, scales1, means1, m1)
    if taps is not None:
        taps["y_hat_0"] = yh0
        taps["spatial_prior"] = sp
    return q0 + q1, (yh0 + yh1) * q_dec, s0 + s1           # inference.py:35-38


# --------------------------------------------------------------------------
# P-frame sub-networks (old: models/video_model.py, refactor: seg_video_model*.py)
# --------------------------------------------------------------------------
def feature_extractor(sd: SD, feature: Tensor, q_feature: Tensor):
    """video_model.py:23-49."""
    x1 = _dcb_chain(sd, "feature_extractor.conv1", 2, feature)
    ctx_t = x1 * q_feature
    ctx = _dcb_chain(sd, "feature_extractor.conv2", 4, x1)
    return ctx, ctx_t


def encoder(sd: SD, variant: str, x_img: Tensor, ctx: Tensor, q_enc: Tensor) -> Tensor:
    """old: video_model.py:52-75; refactor: seg_video_model.py:41-59."""
    f = _conv(sd, "encoder.conv1", F.pixel_unshuffle(x_img, 8))
    f = torch.cat((f, ctx), dim=1)
    if variant == "old":
        f = _dcb_chain(sd, "encoder.conv2", 2, f)
        f = depth_conv_block(sd, "encoder.conv3", f)
    else:
        f = _dcb_chain(sd, "encoder.conv2", 3, f)
    return _conv(sd, "encoder.down", f * q_enc, stride=2, padding=1)


def hyper_encoder(sd: SD, x: Tensor) -> Tensor:
    """video_model.py:123-133."""
    x = depth_conv_block(sd, "hyper_encoder.conv.0", x)
    x = res_block_stride2(sd, "hyper_encoder.conv.1", x)
    return res_block_stride2(sd, "hyper_encoder.conv.2", x)


def hyper_decoder(sd: SD, z_hat: Tensor) -> Tensor:
    """video_model.py:136-146."""
    x = res_block_upsample(sd, "hyper_decoder.conv.0", z_hat)
    x = res_block_upsample(sd, "hyper_decoder.conv.1", x)
    return depth_conv_block(sd, "hyper_decoder.conv.2", x)


def prior_params(sd: SD, z_hat: Tensor, ctx_t: Tensor, taps: Optional[dict] = None) -> Tensor:
    """video_model.py:236-243 (res_prior_param_decoder) + PriorFusion :149-160."""
    hier = hyper_decoder(sd, z_hat)
    temporal = res_block_stride2(sd, "temporal_prior_encoder", ctx_t)
    h, w = temporal.shape[-2:]
    hier = hier[:, :, :h, :w].contiguous()
    if taps is not None:
        taps["hier"] = hier
        taps["temporal"] = temporal
    x = _dcb_chain(sd, "y_prior_fusion.conv", 3, torch.cat((hier, temporal), dim=1))
    return _conv(sd, "y_prior_fusion.conv.3", x)


def spatial_prior(sd: SD, x: Tensor) -> Tensor:
    """video_model.py:163-173."""
    x = _dcb_chain(sd, "y_spatial_prior.conv", 2, x)
    return _conv(sd, "y_spatial_prior.conv.2", x)


def decoder(sd: SD, variant: str, y_hat: Tensor, ctx: Tensor, q_dec: Tensor) -> Tensor:
    """old: video_model.py:78-97 (scale after the last 1x1);
    refactor: seg_video_model.py:62-77 (scale right after `up`)."""
    f = F.pixel_shuffle(_conv(sd, "decoder.up.conv.0", y_hat, padding=1), 2)
    if variant == "old":
        f = _dcb_chain(sd, "decoder.conv1", 3, torch.cat((f, ctx), dim=1))
        return _conv(sd, "decoder.conv2", f) * q_dec
    f = f * q_dec
    f = _dcb_chain(sd, "decoder.conv", 3, torch.cat((f, ctx), dim=1))
    return _conv(sd, "decoder.proj", f)


def recon_generation(sd: SD, feature: Tensor, q_recon: Tensor) -> Tensor:
    """video_model.py:100-120."""
    x = _dcb_chain(sd, "recon_generation_net.conv", 4, feature)
    x = _conv(sd, "recon_generation_net.head", x * q_recon)
    return torch.clamp(F.pixel_shuffle(x, 8), 0.0, 1.0)


def mask_sft(sd: SD, mask_img: Tensor, q_sft: Tensor):
    """performance: seg_video_model.py:159-196."""
    x = _conv(sd, "mask_sft.conv1", F.pixel_unshuffle(mask_img, 8))
    x = _dcb_chain(sd, "mask_sft.conv2", 3, x)
    x = _conv(sd, "mask_sft.down", x * q_sft, stride=2, padding=1)
    return x.chunk(2, dim=1)


def mask_film_hyper_input(sd: SD, y: Tensor, mask_img: Optional[Tensor]) -> Tensor:
    """fast / mask_prop: seg_video_model_fast.py:159-180 (MaskFiLM) and
    :287-325 (_prepare_hyper_input)."""
    b, _, hy, wy = y.shape
    y_pad = pad_for_y(y)
    if mask_img is None:
        m = torch.zeros(b, 1, hy, wy, dtype=y.dtype)
    else:
        m = F.adaptive_avg_pool2d(mask_img.to(y.dtype), (hy, wy)).clamp(0.0, 1.0)
    pb, pr = y_pad.shape[-2] - hy, y_pad.shape[-1] - wy
    if pb or pr:
        m = F.pad(m, (0, pr, 0, pb), mode="constant", value=0.0)
    gb = _conv(sd, "mask_film.net.2", F.relu(_conv(sd, "mask_film.net.0", m, padding=1)))
    gamma, beta = gb.chunk(2, dim=1)
    return y_pad * (1.0 + gamma) + beta


def mask_predictor(sd: SD, prev_mask: Optional[Tensor], ctx: Tensor, ctx_t: Tensor):
    """mask_predictor.py:27-46."""
    if prev_mask is None:
        return None
    hm, wm = prev_mask.shape[-2:]
    hf, wf = ctx.shape[-2:]
    m = F.interpolate(prev_mask, size=(hf, wf), mode="bilinear", align_corners=False)
    m = _conv(sd, "mask_predictor.mask_embed", m, padding=1)
    x = torch.cat([m, ctx, ctx_t], dim=1)
    x = wsilu(_conv(sd, "mask_predictor.net.0", x, padding=1))
    x = wsilu(_conv(sd, "mask_predictor.net.2", x, padding=1))
    logits = _conv(sd, "mask_predictor.net.4", x)
    if (hf, wf) != (hm, wm):
        logits = F.interpolate(logits, size=(hm, wm), mode="bilinear", align_corners=False)
    return logits


# --------------------------------------------------------------------------
# whole-frame forwards
# --------------------------------------------------------------------------
def shift_qp(qp: int, fa_idx: int) -> int:
    """video_model.py:335-336."""
    return qp + QP_SHIFT[fa_idx]


@torch.no_grad()
def dmc_forward(sd: SD, variant: str, x: Tensor, qp: int, dpb: dict, after_i: bool = True,
                taps: Optional[dict] = None) -> dict:
    """DMC.forward: old video_model.py:338-388; performance seg_video_model.py:301-365;
    fast seg_video_model_fast.py:328-411; mask_prop mask_prop_seg_video_model.py:331-417.
    `taps`, when given, receives the intermediate tensors (test instrumentation)."""
    assert variant in VARIANTS
    t = taps if taps is not None else {}
    if variant == "old":
        x_img, mask_img = x, None
    elif x.size(1) > 3:
        x_img, mask_img = x[:, :3], x[:, 3:4]
    else:
        x_img = x
        mask_img = torch.zeros_like(x[:, :1]) if variant == "performance" else None

    q_enc = sd["q_encoder"][qp:qp + 1]
    q_dec = sd["q_decoder"][qp:qp + 1]
    q_feat = sd["q_feature"][qp:qp + 1]
    q_rec = sd["q_recon"][qp:qp + 1]

    if after_i:
        feature = depth_conv_block(sd, "feature_adaptor_i", F.pixel_unshuffle(dpb["frame"], 8))
    else:
        feature = _conv(sd, "feature_adaptor_p", dpb["feature"])
    t["feature_in"] = feature
    ctx, ctx_t = feature_extractor(sd, feature, q_feat)
    t["ctx"], t["ctx_t"] = ctx, ctx_t
    y = encoder(sd, variant, x_img, ctx, q_enc)
    t["y_enc"] = y

    mask_used = None
    if variant == "old":
        hyper_in = pad_for_y(y)
    elif variant == "performance":
        gamma, beta = mask_sft(sd, mask_img, sd["q_sft"][qp:qp + 1])
        t["gamma"], t["beta"] = gamma, beta
        y = y * (1.0 + gamma) + beta
        hyper_in = y                                        # seg_video_model.py:331 (no pad)
    else:
        mask_used = mask_img
        if variant == "mask_prop" and not after_i:
            mask_used = mask_predictor(sd, mask_img, ctx, ctx_t)
        hyper_in = mask_film_hyper_input(sd, y, mask_used)
    t["y"] = y
    t["hyper_in"] = hyper_in

    z = hyper_encoder(sd, hyper_in)
    z_hat = torch.round(z)                                 # inference.py:16-27, eval
    t["z"], t["z_hat"] = z, z_hat
    params = prior_params(sd, z_hat, ctx_t, t)
    t["params"] = params
    sym, y_hat, scales_hat = compress_prior_2x(sd, y, params, t)
    t["y_hat"], t["scales_hat"] = y_hat, scales_hat

    feature = decoder(sd, variant, y_hat, ctx, q_dec)
    x_hat = recon_generation(sd, feature, q_rec)

    pixel_num = x_img.shape[2] * x_img.shape[3]
    if variant == "old":
        bits_y = gaussian_bits_old(sym, scales_hat)
    else:
        sym = sym.clamp(-6.0, 6.0)                         # seg_video_model.py:347 (in place)
        bits_y = gaussian_bits_refactor(sym, scales_hat)
    t["y_q"] = sym
    bits_z = z_bits(sd, z_hat, qp)
    bpp_y = torch.sum(bits_y, dim=(1, 2, 3)) / pixel_num
    bpp_z = torch.sum(bits_z, dim=(1, 2, 3)) / pixel_num
    out = {"dpb": {"frame": x_hat, "feature": feature}, "bpp": bpp_y + bpp_z,
           "bpp_y": bpp_y, "bpp_z": bpp_z}
    if variant in ("fast", "mask_prop"):
        out["mask_pred"] = mask_used if not after_i else None
    return out


def _intra_spatial_prior(sd: SD, x: Tensor) -> Tensor:
    x = _dcb_chain(sd, "y_spatial_prior", 3, x)
    return _conv(sd, "y_spatial_prior.3", x)


@torch.no_grad()
def dmci_forward(sd: SD, x: Tensor, qp: int, taps: Optional[dict] = None) -> dict:
    """DMCI.forward image_model.py:205-261; IntraEncoder :16-43; IntraDecoder :46-93;
    compress_prior_4x models/common_model.py:188-248; separate_prior :171-181."""
    t = taps if taps is not None else {}
    q_enc = sd["q_scale_enc"][qp:qp + 1]
    q_dec = sd["q_scale_dec"][qp:qp + 1]

    f = depth_conv_block(sd, "enc.enc_1", F.pixel_unshuffle(x, 8)) * q_enc
    f = _dcb_chain(sd, "enc.enc_2", 6, f)
    y = _conv(sd, "enc.enc_2.6", f, stride=2, padding=1)
    t["y"] = y

    z = depth_conv_block(sd, "hyper_enc.0", pad_for_y(y))
    z = res_block_stride2(sd, "hyper_enc.1", z)
    z = res_block_stride2(sd, "hyper_enc.2", z)
    z_hat = torch.round(z)
    t["z"], t["z_hat"] = z, z_hat

    p = res_block_upsample(sd, "hyper_dec.0", z_hat)
    p = res_block_upsample(sd, "hyper_dec.1", p)
    p = depth_conv_block(sd, "hyper_dec.2", p)
    p = _dcb_chain(sd, "y_prior_fusion", 3, p)
    params = _conv(sd, "y_prior_fusion.3", p)
    params = params[:, :, :y.shape[2], :y.shape[3]].contiguous()
    t["params"] = params

    # compress_prior_4x
    qq = torch.sigmoid(params[:, :2]) * 1.5 + 0.5
    qe, qd = qq.chunk(2, 1)
    scales, means = params[:, 2:].chunk(2, 1)
    common = _conv(sd, "y_spatial_prior_reduction", params)
    masks = checkerboard_4x(y.shape[1], y.shape[2], y.shape[3], y.dtype)
    ys = y * qe
    q0, yh, s0 = _process_with_mask(ys, scales, means, masks[0])
    q_parts, s_parts = [q0], [s0]
    for step in (1, 2, 3):
        a = depth_conv_block(sd, f"y_spatial_prior_adaptor_{step}", torch.cat((yh, common), dim=1))
        sc, mu = _intra_spatial_prior(sd, a).chunk(2, 1)
        qk, yk, sk = _process_with_mask(ys, sc, mu, masks[step])
        yh = yh + yk
        q_parts.append(qk)
        s_parts.append(sk)
    sym = (q_parts[0] + q_parts[1]) + (q_parts[2] + q_parts[3])
    scales_hat = (s_parts[0] + s_parts[1]) + (s_parts[2] + s_parts[3])
    y_hat = yh * qd
    t["y_q"], t["scales_hat"], t["y_hat"] = sym, scales_hat, y_hat

    d = res_block_upsample(sd, "dec.dec_1.0", y_hat)
    d = _dcb_chain(sd, "dec.dec_1", 12, d, start=1)
    d = depth_conv_block(sd, "dec.dec_2", d * q_dec)
    x_hat = F.pixel_shuffle(d, 8).clamp(0, 1)

    pixel_num = x.shape[2] * x.shape[3]
    bits_y = gaussian_bits_old(sym, scales_hat)
    bits_z = z_bits(sd, z_hat, qp)
    bpp_y = torch.sum(bits_y, dim=(1, 2, 3)) / pixel_num
    bpp_z = torch.sum(bits_z, dim=(1, 2, 3)) / pixel_num
    return {"dpb": {"frame": x_hat, "feature": None}, "bpp": bpp_y + bpp_z,
            "bpp_y": bpp_y, "bpp_z": bpp_z, "bits_y": bits_y.shape, "bits_z": bits_z.shape}


# --------------------------------------------------------------------------
# caller-side metrics (trainer_seg_video_model.py)
# --------------------------------------------------------------------------
def mse(pred: Tensor, target: Tensor) -> Tensor:
    return F.mse_loss(pred, target, reduction="mean")


def roi_mse(pred: Tensor, target: Tensor, mask: Optional[Tensor]) -> Tensor:
    """trainer:655-660: sum(m*(p-t)^2)/sum(m), m=(mask>0) over 3 channels; plain MSE if empty."""
    if mask is None or mask.sum() == 0:
        return mse(pred, target)
    m = (mask > 0).float().expand_as(pred)
    return torch.sum(torch.pow(pred - target, 2) * m) / torch.sum(m)


def psnr_from_mse(m: Tensor) -> Tensor:
    """trainer:598-601."""
    return 10.0 * torch.log10(torch.tensor(1.0, dtype=m.dtype) / (m + 1e-12))


def run_gop(sd_i: SD, sd_p: SD, variant: str, frames: Tensor, masks: Optional[Tensor], qp: int,
            mask_feedback: bool = False):
    """GOP loop of trainer validation_step (trainer:1228-1244): frame 0 through DMCI,
    frames 1.. through the P model with qp shifted by INDEX_MAP.  frames (B,T,3,H,W),
    masks (B,T,1,H,W) or None.  `mask_feedback` applies the SURVEY 8(d) config-4 protocol
    for mask_prop (thresholded mask_pred fed to the next frame)."""
    res = dmci_forward(sd_i, frames[:, 0], qp)
    dpb = res["dpb"]
    outs = [res]
    prev_pred = None
    for ti in range(1, frames.shape[1]):
        cq = shift_qp(qp, INDEX_MAP[ti % 8])
        if variant == "old" or masks is None:
            x_in = frames[:, ti]
        else:
            m = masks[:, ti]
            if mask_feedback and variant == "mask_prop":
                if ti == 2:
                    m = masks[:, 1]
                elif ti >= 3 and prev_pred is not None:
                    m = (prev_pred > 0).float()
            x_in = torch.cat([frames[:, ti], m], dim=1)
        res = dmc_forward(sd_p, variant, x_in, cq, dpb, after_i=(ti == 1))
        prev_pred = res.get("mask_pred")
        dpb = res["dpb"]
        outs.append(res)
    return outs
