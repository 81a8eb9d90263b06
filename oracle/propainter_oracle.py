"""CPU restatement of the ProPainter inference hot path (the parity oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Nothing in the product package imports it.

Why it exists: the reference (daniabib/ComfyUI_ProPainter_Nodes) is a Python package that cannot
travel to the GPU box, and it ships no tests or golden vectors ("parity unpinned" by the reference
itself).  This file restates the algorithm functionally -- plain functions over a ``state_dict``,
no nn.Module -- and is *pinned* against outputs of the real reference generated in the build
container (tests/golden/, produced by tests/golden/make_golden.py, checked by
tests/test_oracle_golden.py).  The arithmetic lives in the same third-party libraries the
reference calls (PyTorch conv/grid_sample/fold/unfold, torchvision.ops.deform_conv2d).

Every function cites the reference lines it follows (paths relative to the reference repo).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torchvision.ops import deform_conv2d

# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------


def _c2(sd, name, x, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride, padding, dilation, groups)


def _lrelu(x, s):
    return F.leaky_relu(x, s)


def strip_module_prefix(sd):
    """RAFT checkpoints carry a DataParallel ``module.`` prefix (model/modules/flow_comp_raft.py:17-19)."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def warp_by_flow(x, flow_nhw2, mode="bilinear"):
    """flow_warp (model/modules/flow_loss_utils.py:6-51): sample x at (pixel + flow), zeros outside,
    align_corners=True; the grid is built in x's dtype."""
    _, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    base = torch.stack((gx, gy), 2).to(x.dtype)
    g = base + flow_nhw2
    nx = 2.0 * g[..., 0] / max(w - 1, 1) - 1.0
    ny = 2.0 * g[..., 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((nx, ny), 3), mode=mode, padding_mode="zeros", align_corners=True)


def fb_consistency(flow_fw, flow_bw, a1=0.01, a2=0.5):
    """fbConsistencyCheck (model/propainter.py:27-36)."""
    bw_w = warp_by_flow(flow_bw, flow_fw.permute(0, 2, 3, 1))
    diff = flow_fw + bw_w
    mag = (flow_fw ** 2).sum(1, keepdim=True) + (bw_w ** 2).sum(1, keepdim=True)
    return ((diff ** 2).sum(1, keepdim=True) < a1 * mag + a2).to(flow_fw.dtype)


# ----------------------------------------------------------------------------------------------
# stage 1: RAFT (bidirectional)
# ----------------------------------------------------------------------------------------------


def _norm(sd, name, x, kind):
    if kind == "instance":  # InstanceNorm2d, affine=False, per-sample statistics
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, 1e-5)


def raft_encoder(sd, p, x, kind):
    """BasicEncoder (model/modules/RAFT/extractor.py:121-193), ResidualBlock (:5-57)."""
    x = F.relu(_norm(sd, p + "norm1", _c2(sd, p + "conv1", x, 2, 3), kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        for bi in range(2):
            q = f"{p}layer{li}.{bi}."
            s = stride if bi == 0 else 1
            y = F.relu(_norm(sd, q + "norm1", _c2(sd, q + "conv1", x, s, 1), kind))
            y = F.relu(_norm(sd, q + "norm2", _c2(sd, q + "conv2", y, 1, 1), kind))
            if s != 1:
                x = _norm(sd, q + "norm3", _c2(sd, q + "downsample.0", x, s, 0), kind)
            x = F.relu(x + y)
    return _c2(sd, p + "conv2", x)


def corr_pyramid(f1, f2, levels=4):
    """CorrBlock.__init__/corr (model/modules/RAFT/corr.py:13-27,52-60)."""
    b, d, h, w = f1.shape
    c = torch.matmul(f1.view(b, d, h * w).transpose(1, 2), f2.view(b, d, h * w)) / math.sqrt(d)
    c = c.reshape(b * h * w, 1, h, w)
    pyr = [c]
    for _ in range(levels - 1):
        c = F.avg_pool2d(c, 2, stride=2)
        pyr.append(c)
    return pyr


def corr_lookup(pyr, coords, r=4):
    """CorrBlock.__call__ (corr.py:29-50) + bilinear_sampler (RAFT/utils/utils.py:66-80).

    Channel ``l*81 + i*9 + j`` samples level l at (x/2^l + (i-4), y/2^l + (j-4)): the first window
    index moves x (the meshgrid(dy, dx) quirk at corr.py:37-39)."""
    b, _, h, w = coords.shape
    co = coords.permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2)
    d = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
    out = []
    for lvl, c in enumerate(pyr):
        pts = co / 2 ** lvl + delta
        hh, ww = c.shape[-2:]
        gx = 2 * pts[..., 0:1] / (ww - 1) - 1
        gy = 2 * pts[..., 1:2] / (hh - 1) - 1
        s = F.grid_sample(c, torch.cat([gx, gy], -1), align_corners=True)
        out.append(s.view(b, h, w, -1))
    return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous().float()


def raft_update(sd, net, inp, corr, flow):
    """BasicUpdateBlock (model/modules/RAFT/update.py:94-154)."""
    u = "update_block."
    cor = F.relu(_c2(sd, u + "encoder.convc1", corr))
    cor = F.relu(_c2(sd, u + "encoder.convc2", cor, 1, 1))
    flo = F.relu(_c2(sd, u + "encoder.convf1", flow, 1, 3))
    flo = F.relu(_c2(sd, u + "encoder.convf2", flo, 1, 1))
    mot = torch.cat([F.relu(_c2(sd, u + "encoder.conv", torch.cat([cor, flo], 1), 1, 1)), flow], 1)
    x = torch.cat([inp, mot], 1)
    for sfx, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([net, x], 1)
        z = torch.sigmoid(_c2(sd, u + "gru.convz" + sfx, hx, 1, pad))
        r = torch.sigmoid(_c2(sd, u + "gru.convr" + sfx, hx, 1, pad))
        q = torch.tanh(_c2(sd, u + "gru.convq" + sfx, torch.cat([r * net, x], 1), 1, pad))
        net = (1 - z) * net + z * q
    delta = _c2(sd, u + "flow_head.conv2", F.relu(_c2(sd, u + "flow_head.conv1", net, 1, 1)), 1, 1)
    mask = 0.25 * _c2(sd, u + "mask.2", F.relu(_c2(sd, u + "mask.0", net, 1, 1)))
    return net, mask, delta


def convex_upsample(flow, mask):
    """RAFT.upsample_flow (model/modules/RAFT/raft.py:81-92)."""
    n, _, h, w = flow.shape
    m = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, 8 * h, 8 * w)


def raft_pairs(sd, img1, img2, iters, return_trace=False):
    """RAFT.forward in test mode (raft.py:94-152); inputs already in [-1, 1] (:96-97 commented out)."""
    n, _, H, W = img1.shape
    f = raft_encoder(sd, "fnet.", torch.cat([img1, img2], 0), "instance").float()
    f1, f2 = f[:n], f[n:]
    pyr = corr_pyramid(f1, f2)
    c = raft_encoder(sd, "cnet.", img1, "batch")
    net, inp = torch.tanh(c[:, :128]), torch.relu(c[:, 128:])
    ys, xs = torch.meshgrid(torch.arange(H // 8), torch.arange(W // 8), indexing="ij")
    coords0 = torch.stack([xs, ys], 0).float()[None].repeat(n, 1, 1, 1)
    coords1 = coords0.clone()
    trace = []
    mask = None
    for _ in range(iters):
        corr = corr_lookup(pyr, coords1)
        net, mask, delta = raft_update(sd, net, inp, corr, coords1 - coords0)
        coords1 = coords1 + delta
        if return_trace:   # raft.py:141-147: flow_predictions holds the up-sampled flow of every iteration
            trace.append(convex_upsample(coords1 - coords0, mask))
    up = convex_upsample(coords1 - coords0, mask)
    return (up, trace) if return_trace else up


def raft_bidirectional(sd, frames, iters=20):
    """RAFT_bi.forward (model/modules/flow_comp_raft.py:39-58). frames [1,l,3,H,W] -> 2x[1,l-1,2,H,W]."""
    sd = strip_module_prefix(sd)
    b, l, c, h, w = frames.shape
    a = frames[:, :-1].reshape(-1, c, h, w)
    bb = frames[:, 1:].reshape(-1, c, h, w)
    ff = raft_pairs(sd, a, bb, iters)
    fb = raft_pairs(sd, bb, a, iters)
    return ff.view(b, l - 1, 2, h, w), fb.view(b, l - 1, 2, h, w)


def raft_clip_length(width):
    """propainter_inference.py:65-72."""
    if width <= 640:
        return 12
    if width <= 720:
        return 8
    if width <= 1280:
        return 4
    return 2


def compute_flow(raft_sd, frames, raft_iter):
    """compute_flow (propainter_inference.py:61-99): clips of <= short_clip_len with one-frame overlap."""
    T = frames.shape[1]
    clip = raft_clip_length(frames.shape[-1])
    if T <= clip:
        return raft_bidirectional(raft_sd, frames, raft_iter)
    ff, fb = [], []
    for s in range(0, T, clip):
        e = min(T, s + clip)
        a, b = raft_bidirectional(raft_sd, frames[:, (s if s == 0 else s - 1):e], raft_iter)
        ff.append(a)
        fb.append(b)
    return torch.cat(ff, 1), torch.cat(fb, 1)


# ----------------------------------------------------------------------------------------------
# stage 2: recurrent flow completion
# ----------------------------------------------------------------------------------------------


def _c3(sd, name, x, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1)):
    return F.conv3d(x, sd[name + ".weight"], sd[name + ".bias"], stride, padding, dilation)


def _p3d(sd, name, x, stride):
    """P3DBlock (model/recurrent_flow_completion.py:162-205): (1,3,3) conv + LReLU + (3,1,1) dil-2 conv."""
    y = _lrelu(_c3(sd, name + ".conv1.0", x, (1, stride, stride), (0, 1, 1)), 0.2)
    return _c3(sd, name + ".conv2.0", y, (1, 1, 1), (2, 0, 0), (2, 1, 1))


def _deform_align2(sd, p, x, cond, max_mag=5.0):
    """SecondOrderDeformableAlignment.forward (recurrent_flow_completion.py:32-53)."""
    o = cond
    for i in (0, 2, 4):
        o = _lrelu(_c2(sd, f"{p}.conv_offset.{i}", o, 1, 1), 0.1)
    o = _c2(sd, p + ".conv_offset.6", o, 1, 1)
    o1, o2, m = torch.chunk(o, 3, dim=1)
    offset = max_mag * torch.tanh(torch.cat((o1, o2), 1))
    return deform_conv2d(x, offset, sd[p + ".weight"], sd[p + ".bias"], 1, 1, 1, torch.sigmoid(m))


def rfc_feat_prop(sd, x):
    """BidirectionalPropagation.forward of the flow net (recurrent_flow_completion.py:77-143). x [b,t,c,h,w]."""
    b, t, c, h, w = x.shape
    fp = "feat_prop_module."
    spatial = [x[:, i] for i in range(t)]
    feats = {}
    for name in ("backward_", "forward_"):
        order = list(range(t))[::-1] if name == "backward_" else list(range(t))
        outs = []
        prop = x.new_zeros(b, c, h, w)
        for i, idx in enumerate(order):
            cur = spatial[idx]
            if i > 0:
                n2 = outs[-2] if i > 1 else torch.zeros_like(prop)
                cond = torch.cat([prop, cur, n2], 1)
                prop = _deform_align2(sd, fp + "deform_align." + name, torch.cat([prop, n2], 1), cond)
            parts = [cur]
            if name == "forward_":
                parts.append(feats["backward_"][idx])
            parts.append(prop)
            y = _lrelu(_c2(sd, fp + f"backbone.{name}.0", torch.cat(parts, 1), 1, 1), 0.1)
            prop = prop + _c2(sd, fp + f"backbone.{name}.2", y, 1, 1)
            outs.append(prop)
        feats[name] = outs[::-1] if name == "backward_" else outs
    out = [_c2(sd, fp + "fusion", torch.cat([feats["backward_"][i], feats["forward_"][i]], 1)) for i in range(t)]
    return torch.stack(out, 1) + x


def _deconv(sd, name, x):
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return _c2(sd, name + ".conv", x, 1, 1)


def rfc_forward(sd, masked_flows, masks):
    """RecurrentFlowCompleteNet.forward in eval (recurrent_flow_completion.py:315-354)."""
    b, t, _, h, w = masked_flows.shape
    inp = torch.cat((masked_flows.permute(0, 2, 1, 3, 4), masks.permute(0, 2, 1, 3, 4)), 1)
    xp = F.pad(inp, (2, 2, 2, 2, 0, 0), mode="replicate")  # padding_mode="replicate", pad (0,2,2)
    x = _lrelu(_c3(sd, "downsample.0", xp, (1, 2, 2)), 0.2)
    e1 = _lrelu(_p3d(sd, "encoder1.2", _lrelu(_p3d(sd, "encoder1.0", x, 1), 0.2), 2), 0.2)
    e2 = _lrelu(_p3d(sd, "encoder2.2", _lrelu(_p3d(sd, "encoder2.0", e1, 1), 0.2), 2), 0.2)
    m = e2
    for i, d in ((0, 3), (2, 2), (4, 1)):
        m = _lrelu(_c3(sd, f"mid_dilation.{i}", m, (1, 1, 1), (0, d, d), (1, d, d)), 0.2)
    prop = rfc_feat_prop(sd, m.permute(0, 2, 1, 3, 4)).reshape(-1, 128, h // 8, w // 8)
    e1f = e1.permute(0, 2, 1, 3, 4).reshape(-1, e1.shape[1], e1.shape[3], e1.shape[4])
    d2 = _lrelu(_deconv(sd, "decoder2.2", _lrelu(_c2(sd, "decoder2.0", prop, 1, 1), 0.2)), 0.2) + e1f
    d1 = _lrelu(_deconv(sd, "decoder1.2", _lrelu(_c2(sd, "decoder1.0", d2, 1, 1), 0.2)), 0.2)
    flow = _deconv(sd, "upsample.2", _lrelu(_c2(sd, "upsample.0", d1, 1, 1), 0.2))
    return flow.view(b, t, 2, h, w)


def rfc_bidirectional(sd, flows_bi, masks):
    """forward_bidirect_flow + combine_flow (recurrent_flow_completion.py:356-400)."""
    mf, mb = masks[:, :-1], masks[:, 1:]
    pf = rfc_forward(sd, flows_bi[0] * (1 - mf), mf)
    pb = rfc_forward(sd, torch.flip(flows_bi[1] * (1 - mb), [1]), torch.flip(mb, [1]))
    pb = torch.flip(pb, [1])
    return pf * mf + flows_bi[0] * (1 - mf), pb * mb + flows_bi[1] * (1 - mb)


def complete_flow(sd, flows_bi, flow_masks, subvideo_length):
    """complete_flow (propainter_inference.py:102-156): chunks of subvideo_length with 5-flow halo."""
    L = flows_bi[0].shape[1]
    if L <= subvideo_length:
        return rfc_bidirectional(sd, flows_bi, flow_masks)
    pad = 5
    of, ob = [], []
    for f in range(0, L, subvideo_length):
        s, e = max(0, f - pad), min(L, f + subvideo_length + pad)
        ps, pe = f - s, e - min(L, f + subvideo_length)
        a, b = rfc_bidirectional(sd, (flows_bi[0][:, s:e], flows_bi[1][:, s:e]), flow_masks[:, s:e + 1])
        of.append(a[:, ps:e - s - pe])
        ob.append(b[:, ps:e - s - pe])
    return torch.cat(of, 1), torch.cat(ob, 1)


# ----------------------------------------------------------------------------------------------
# stage 3a: image propagation (non-learnable)
# ----------------------------------------------------------------------------------------------


def _bin(m, th=0.1):
    return (m > th).to(m.dtype)


def img_propagation(masked_frames, flows_f, flows_b, masks, interpolation="nearest"):
    """BidirectionalPropagation(3, learnable=False).forward (model/propainter.py:118-231).

    Returns (frames after backward+forward passes [b,t,c,h,w], masks after the forward pass)."""
    b, t, c, h, w = masked_frames.shape
    feats = [masked_frames[:, i] for i in range(t)]
    msks = [masks[:, i] for i in range(t)]
    for name in ("backward", "forward"):
        if name == "backward":
            order = list(range(t))[::-1]
            fidx = order
            f_prop, f_chk = flows_f, flows_b
        else:
            order = list(range(t))
            fidx = list(range(-1, t - 1))
            f_prop, f_chk = flows_b, flows_f
        of, om = [], []
        for i, idx in enumerate(order):
            cur, mcur = feats[idx], msks[idx]
            if i == 0:
                prop, mprop = cur, mcur
            else:
                fp_, fc_ = f_prop[:, fidx[i]], f_chk[:, fidx[i]]
                valid = fb_consistency(fp_, fc_)
                warped = warp_by_flow(prop, fp_.permute(0, 2, 3, 1), interpolation)
                mv = _bin(warp_by_flow(mprop, fp_.permute(0, 2, 3, 1)))
                u = _bin(mcur * valid * (1 - mv))
                prop = u * warped + (1 - u) * cur
                mprop = _bin(mcur * (1 - (valid * (1 - mv))))
            of.append(prop)
            om.append(mprop)
        if name == "backward":
            of, om = of[::-1], om[::-1]
        feats, msks = of, om
    return torch.stack(feats, 1), torch.stack(msks, 1)


def image_propagation(frames, masks_dilated, flows_bi, subvideo_length):
    """image_propagation (propainter_inference.py:159-225)."""
    T = frames.shape[1]
    masked = frames * (1 - masks_dilated)
    sub = min(100, subvideo_length)
    if T <= sub:
        p, m = img_propagation(masked, flows_bi[0], flows_bi[1], masks_dilated)
        return frames * (1 - masks_dilated) + p * masks_dilated, m
    pad = 10
    uf, um = [], []
    for f in range(0, T, sub):
        s, e = max(0, f - pad), min(T, f + sub + pad)
        ps, pe = f - s, e - min(T, f + sub)
        p, m = img_propagation(masked[:, s:e], flows_bi[0][:, s:e - 1], flows_bi[1][:, s:e - 1], masks_dilated[:, s:e])
        u = frames[:, s:e] * (1 - masks_dilated[:, s:e]) + p * masks_dilated[:, s:e]
        uf.append(u[:, ps:e - s - pe])
        um.append(m[:, ps:e - s - pe])
    return torch.cat(uf, 1), torch.cat(um, 1)


# ----------------------------------------------------------------------------------------------
# stage 3b: InpaintGenerator (encoder, feature propagation, sparse transformer, decoder)
# ----------------------------------------------------------------------------------------------

ENC_GROUPS = {10: 2, 12: 4, 14: 8, 16: 1}


def gen_encoder(sd, x):
    """Encoder.forward (model/propainter.py:234-275): grouped convs see cat(x0 slice, out slice) per group."""
    bt = x.shape[0]
    out = x
    x0 = None
    for idx, stride in ((0, 2), (2, 1), (4, 2), (6, 1), (8, 1), (10, 1), (12, 1), (14, 1), (16, 1)):
        g = 1
        if idx == 8:
            x0 = out
        if idx > 8:
            g = ENC_GROUPS[idx]
            h, w = x0.shape[-2:]
            out = torch.cat([x0.view(bt, g, -1, h, w), out.view(bt, g, -1, h, w)], 2).view(bt, -1, h, w)
        out = _lrelu(_c2(sd, f"encoder.layers.{idx}", out, stride, 1, 1, g), 0.2)
    return out


def _deform_align_flow(sd, p, x, cond, flow, max_mag=3.0):
    """DeformableAlignment.forward (model/propainter.py:62-82)."""
    o = cond
    for i in (0, 2, 4):
        o = _lrelu(_c2(sd, f"{p}.conv_offset.{i}", o, 1, 1), 0.1)
    o = _c2(sd, p + ".conv_offset.6", o, 1, 1)
    o1, o2, m = torch.chunk(o, 3, dim=1)
    offset = max_mag * torch.tanh(torch.cat((o1, o2), 1))
    offset = offset + flow.flip(1).repeat(1, offset.size(1) // 2, 1, 1)
    return deform_conv2d(x, offset, sd[p + ".weight"], sd[p + ".bias"], 1, 1, 1, torch.sigmoid(m))


def gen_feat_prop(sd, x, flows_f, flows_b, mask2):
    """BidirectionalPropagation(128, learnable=True).forward (model/propainter.py:118-231)."""
    b, t, c, h, w = x.shape
    fp = "feat_prop_module."
    feats = [x[:, i] for i in range(t)]
    msk = [mask2[:, i] for i in range(t)]
    outs = {}
    for name in ("backward_1", "forward_1"):
        if name == "backward_1":
            order = list(range(t))[::-1]
            fidx = order
            f_prop, f_chk = flows_f, flows_b
        else:
            order = list(range(t))
            fidx = list(range(-1, t - 1))
            f_prop, f_chk = flows_b, flows_f
        res = []
        for i, idx in enumerate(order):
            cur, mcur = feats[idx], msk[idx]
            if i == 0:
                prop = cur
            else:
                fl, fc_ = f_prop[:, fidx[i]], f_chk[:, fidx[i]]
                valid = fb_consistency(fl, fc_)
                warped = warp_by_flow(prop, fl.permute(0, 2, 3, 1), "bilinear")
                cond = torch.cat([cur, warped, fl, valid, mcur], 1)
                prop = _deform_align_flow(sd, fp + "deform_align." + name, prop, cond, fl)
            y = _lrelu(_c2(sd, fp + f"backbone.{name}.0", torch.cat([cur, prop, mcur], 1), 1, 1), 0.2)
            prop = prop + _c2(sd, fp + f"backbone.{name}.2", y, 1, 1)
            res.append(prop)
        if name == "backward_1":
            res = res[::-1]
        outs[name] = res
        feats = res  # the forward pass consumes the backward pass's outputs (propainter.py:131,150-151)
    ob = torch.stack(outs["backward_1"], 1).view(-1, c, h, w)
    of = torch.stack(outs["forward_1"], 1).view(-1, c, h, w)
    y = _lrelu(_c2(sd, fp + "fuse.0", torch.cat([ob, of, mask2.view(-1, 2, h, w)], 1), 1, 1), 0.2)
    return (_c2(sd, fp + "fuse.2", y, 1, 1) + x.view(-1, c, h, w)).view(b, t, c, h, w)


T2T = dict(kernel_size=(7, 7), stride=(3, 3), padding=(3, 3))


def soft_split(sd, x, b):
    """SoftSplit (model/modules/sparse_transformer.py:8-36)."""
    h, w = x.shape[-2:]
    fh = int((h + 2 * 3 - 6 - 1) / 3 + 1)
    fw = int((w + 2 * 3 - 6 - 1) / 3 + 1)
    f = F.unfold(x, **T2T).permute(0, 2, 1)
    f = F.linear(f, sd["ss.embedding.weight"], sd["ss.embedding.bias"])
    return f.view(b, -1, fh, fw, f.size(2))


def soft_comp(sd, x, t, size):
    """SoftComp (sparse_transformer.py:39-64)."""
    b_, _, _, _, c_ = x.shape
    f = F.linear(x.view(b_, -1, c_), sd["sc.embedding.weight"], sd["sc.embedding.bias"])
    f = f.view(b_ * t, -1, f.size(2)).permute(0, 2, 1)
    f = F.fold(f, output_size=size, **T2T)
    return _c2(sd, "sc.bias_conv", f, 1, 1)


def fusion_ffn(sd, p, x, size):
    """FusionFeedForward (sparse_transformer.py:67-123)."""
    n_vecs = 1
    for i, d in enumerate(T2T["kernel_size"]):
        n_vecs *= int((size[i] + 2 * T2T["padding"][i] - (d - 1) - 1) / T2T["stride"][i] + 1)
    x = F.linear(x, sd[p + "fc1.0.weight"], sd[p + "fc1.0.bias"])
    b, n, c = x.shape
    ones = x.new_ones(b, n, 49).view(-1, n_vecs, 49).permute(0, 2, 1)
    norm = F.fold(ones, output_size=size, **T2T)
    y = F.fold(x.view(-1, n_vecs, c).permute(0, 2, 1), output_size=size, **T2T)
    y = F.unfold(y / norm, **T2T).permute(0, 2, 1).contiguous().view(b, n, c)
    return F.linear(F.gelu(y), sd[p + "fc2.1.weight"], sd[p + "fc2.1.bias"])


def _win_part(x, ws, nh):
    B, T, H, W, C = x.shape
    x = x.view(B, T, H // ws[0], ws[0], W // ws[1], ws[1], nh, C // nh)
    return x.permute(0, 2, 4, 6, 1, 3, 5, 7).contiguous()


def sparse_window_attention(sd, p, x, mask, t_ind, ws=(5, 9), nh=4):
    """SparseWindowAttention.forward (sparse_transformer.py:201-393)."""
    b, t, h, w, c = x.shape
    wh, ww = ws
    ch = c // nh
    nwh, nww = math.ceil(h / wh), math.ceil(w / ww)
    nh_, nw_ = nwh * wh, nww * ww
    pr, pb = nw_ - w, nh_ - h
    if pr > 0 or pb > 0:
        x = F.pad(x, (0, 0, 0, pr, 0, pb, 0, 0))
        mask = F.pad(mask, (0, 0, 0, pr, 0, pb, 0, 0))
    lin = lambda n, v: F.linear(v, sd[p + n + ".weight"], sd[p + n + ".bias"])
    q, k, v = lin("query", x), lin("key", x), lin("value", x)
    nW = nwh * nww
    part = lambda a: _win_part(a.contiguous(), ws, nh).view(b, nW, nh, t, wh * ww, ch)
    wq, wk, wv = part(q), part(k), part(v)
    eh, ew = (wh + 1) // 2, (ww + 1) // 2
    valid = sd[p + "valid_ind_rolled"]
    rk, rv = [], []
    for sh in ((-eh, -ew), (-eh, ew), (eh, -ew), (eh, ew)):
        rk.append(part(torch.roll(k, shifts=sh, dims=(2, 3))))
        rv.append(part(torch.roll(v, shifts=sh, dims=(2, 3))))
    wk = torch.cat((wk, torch.cat(rk, 4)[:, :, :, :, valid]), 4)
    wv = torch.cat((wv, torch.cat(rv, 4)[:, :, :, :, valid]), 4)
    px = F.conv2d(x.view(b * t, nh_, nw_, c).permute(0, 3, 1, 2), sd[p + "pool_layer.weight"],
                  sd[p + "pool_layer.bias"], stride=4, groups=c)
    ph, pw = px.shape[-2:]
    px = px.permute(0, 2, 3, 1).view(b, t, ph, pw, c)

    def pooled(n):
        y = lin(n, px).unsqueeze(1).repeat(1, nW, 1, 1, 1, 1)
        y = y.view(b, nW, t, ph, pw, nh, ch).permute(0, 1, 5, 2, 3, 4, 6)
        return y.contiguous().view(b, nW, nh, t, ph * pw, ch)

    wk = torch.cat((wk, pooled("key")), 4)
    wv = torch.cat((wv, pooled("value")), 4)
    out = torch.zeros_like(wq)
    lt = mask.size(1)
    wm = F.max_pool2d(mask.view(b * lt, nh_, nw_), ws, ws).view(b, lt, nW).sum(1)
    for i in range(b):
        mi = wm[i].nonzero(as_tuple=False).view(-1)
        if len(mi) > 0:
            qt = wq[i, mi].view(len(mi), nh, t * wh * ww, ch)
            kt = wk[i, mi][:, :, t_ind.view(-1)].reshape(len(mi), nh, -1, ch)
            vt = wv[i, mi][:, :, t_ind.view(-1)].reshape(len(mi), nh, -1, ch)
            a = F.softmax((qt @ kt.transpose(-2, -1)) * (1.0 / math.sqrt(ch)), dim=-1)
            out[i, mi] = (a @ vt).view(-1, nh, t, wh * ww, ch)
        ui = (wm[i] == 0).nonzero(as_tuple=False).view(-1)
        qs, ks, vs = wq[i, ui], wk[i, ui, :, :, :wh * ww], wv[i, ui, :, :, :wh * ww]
        a = F.softmax((qs @ ks.transpose(-2, -1)) * (1.0 / math.sqrt(ch)), dim=-1)
        out[i, ui] = a @ vs
    out = out.view(b, nwh, nww, nh, t, wh, ww, ch).permute(0, 4, 1, 5, 2, 6, 3, 7).contiguous()
    out = out.view(b, t, nh_, nw_, c)[:, :, :h, :w]
    return lin("proj", out)


def transformer_blocks(sd, x, size, l_mask, depths=8, t_dilation=2):
    """TemporalSparseTransformerBlock / TemporalSparseTransformer (sparse_transformer.py:396-467)."""
    B, T, H, W, C = x.shape
    t_inds = [torch.arange(i, T, t_dilation) for i in range(t_dilation)] * (depths // t_dilation)
    for i in range(depths):
        p = f"transformers.transformer.{i}."
        y = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = x + sparse_window_attention(sd, p + "attention.", y, l_mask, t_inds[i])
        y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        x = x + fusion_ffn(sd, p + "mlp.", y.view(B, T * H * W, C), size).view(B, T, H, W, C)
    return x


def gen_decoder(sd, x):
    """decoder Sequential (model/propainter.py:304-312)."""
    x = _lrelu(_deconv(sd, "decoder.0", x), 0.2)
    x = _lrelu(_c2(sd, "decoder.2", x, 1, 1), 0.2)
    x = _lrelu(_deconv(sd, "decoder.4", x), 0.2)
    return _c2(sd, "decoder.6", x, 1, 1)


def inpaint_window(sd, frames, flows_bi, masks_in, masks_updated, l_t, return_parts=False):
    """InpaintGenerator.forward in eval (model/propainter.py:358-453)."""
    b, t, _, H, W = frames.shape
    enc = gen_encoder(sd, torch.cat([frames.view(b * t, 3, H, W), masks_in.view(b * t, 1, H, W),
                                     masks_updated.view(b * t, 1, H, W)], 1))
    _, c, h, w = enc.shape
    enc = enc.view(b, t, c, h, w)
    local, ref = enc[:, :l_t], enc[:, l_t:]
    ds = lambda f: F.interpolate(f.reshape(-1, 2, H, W), scale_factor=1 / 4, mode="bilinear",
                                 align_corners=False).view(b, l_t - 1, 2, h, w) / 4.0
    dff, dfb = ds(flows_bi[0]), ds(flows_bi[1])
    dmi = F.interpolate(masks_in.reshape(-1, 1, H, W), scale_factor=1 / 4, mode="nearest").view(b, t, 1, h, w)
    dmi_l = dmi[:, :l_t]
    dmu_l = F.interpolate(masks_updated[:, :l_t].reshape(-1, 1, H, W), scale_factor=1 / 4,
                          mode="nearest").view(b, l_t, 1, h, w)
    mp = F.max_pool2d(dmi_l.reshape(-1, 1, h, w), (7, 7), (3, 3), (3, 3))
    mp = mp.view(b, l_t, 1, mp.size(-2), mp.size(-1)).permute(0, 1, 3, 4, 2).contiguous()
    local_p = gen_feat_prop(sd, local, dff, dfb, torch.cat([dmi_l, dmu_l], 2))
    enc2 = torch.cat((local_p, ref), 1)
    tok = soft_split(sd, enc2.view(-1, c, h, w), b)
    tok2 = transformer_blocks(sd, tok, (h, w), mp)
    tr = soft_comp(sd, tok2, t, (h, w)).view(b, t, -1, h, w)
    enc3 = enc2 + tr
    out = torch.tanh(gen_decoder(sd, enc3[:, :l_t].reshape(-1, c, h, w))).view(b, l_t, 3, H, W)
    if return_parts:
        return out, dict(enc=enc, local_prop=local_p, tokens=tok, tokens_out=tok2, enc_out=enc3)
    return out


# ----------------------------------------------------------------------------------------------
# window scheduling + host composite
# ----------------------------------------------------------------------------------------------


def ref_indices(mid, neighbor_ids, video_length, ref_stride, ref_num):
    """get_ref_index (propainter_inference.py:36-58)."""
    out = []
    if ref_num == -1:
        return [i for i in range(0, video_length, ref_stride) if i not in neighbor_ids]
    s = max(0, mid - ref_stride * (ref_num // 2))
    e = min(video_length, mid + ref_stride * (ref_num // 2))
    for i in range(s, e, ref_stride):
        if i not in neighbor_ids:
            if len(out) > ref_num:
                break
            out.append(i)
    return out


def window_schedule(video_length, neighbor_length, ref_stride, subvideo_length):
    """The (neighbor_ids, ref_ids) list walked by feature_propagation (propainter_inference.py:245-262)."""
    stride = neighbor_length // 2
    ref_num = subvideo_length // ref_stride if video_length > subvideo_length else -1
    sched = []
    for f in range(0, video_length, stride):
        nb = list(range(max(0, f - stride), min(video_length, f + stride + 1)))
        sched.append((nb, ref_indices(f, nb, video_length, ref_stride, ref_num)))
    return sched


def composite_window(comp, pred255, binary_masks, original_frames, neighbor_ids):
    """The uint8 composite of one window (propainter_inference.py:294-307), in place on the list ``comp``.

    pred255: [l_t,H,W,3] float array = ((pred+1)/2) * 255 in the dtype the reference holds at that point (float32 with
    fp16="disable"; float16 with fp16="enable", where both the +1, /2 on the device and the numpy *255 round to half);
    binary_masks: [l_t,H,W,1] uint8 dilated masks; original_frames: list of HxWx3 uint8."""
    for i, idx in enumerate(neighbor_ids):
        img = np.array(pred255[i]).astype(np.uint8) * binary_masks[i] + original_frames[idx] * (1 - binary_masks[i])
        if comp[idx] is None:
            comp[idx] = img
        else:
            comp[idx] = comp[idx].astype(np.float32) * 0.5 + img.astype(np.float32) * 0.5
        comp[idx] = comp[idx].astype(np.uint8)


def feature_propagation(gen_sd, updated_frames, updated_masks, masks_dilated, flows_bi, original_frames,
                        neighbor_length, ref_stride, subvideo_length):
    """feature_propagation incl. the host composite (propainter_inference.py:228-311)."""
    T = updated_frames.shape[1]
    H, W = updated_frames.shape[-2:]
    comp = [None] * T
    for nb, refs in window_schedule(T, neighbor_length, ref_stride, subvideo_length):
        ids = nb + refs
        pred = inpaint_window(gen_sd, updated_frames[:, ids], (flows_bi[0][:, nb[:-1]], flows_bi[1][:, nb[:-1]]),
                              masks_dilated[:, ids], updated_masks[:, ids], len(nb))
        pred = ((pred.view(-1, 3, H, W) + 1) / 2).cpu().permute(0, 2, 3, 1).numpy() * 255
        bm = masks_dilated[0, nb].cpu().permute(0, 2, 3, 1).numpy().astype(np.uint8)
        composite_window(comp, pred, bm, original_frames, nb)
    return comp


def run_pipeline(raft_sd, rfc_sd, gen_sd, frames, flow_masks, masks_dilated, original_frames, *,
                 raft_iter=20, subvideo_length=80, neighbor_length=10, ref_stride=10, return_stages=False):
    """process_inpainting + feature_propagation (propainter_inference.py:314-341, 228-311), fp32 CPU."""
    with torch.no_grad():
        gt = compute_flow(raft_sd, frames, raft_iter)
        pred = complete_flow(rfc_sd, gt, flow_masks, subvideo_length)
        uf, um = image_propagation(frames, masks_dilated, pred, subvideo_length)
        comp = feature_propagation(gen_sd, uf, um, masks_dilated, pred, original_frames,
                                   neighbor_length, ref_stride, subvideo_length)
    if return_stages:
        return comp, dict(gt_flows=gt, pred_flows=pred, updated_frames=uf, updated_masks=um)
    return comp
