"""ctypes binding of libpropainter_b200.so + checkpoint packing for the sm_100a kernels.

PyTorch is used here only for device memory, streams and host-side weight re-layout; every compute
step goes through the C ABI declared in include/propainter_b200.h.  There is NO fallback path: if the
shared library is missing or the device is not a B200, construction fails loudly.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Tuple

import torch

from . import weights as Wspec

_LIB = None
_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PP_LIB_PATH") or os.path.join(_PKG_DIR, "libpropainter_b200.so")   # PP_LIB_PATH: A/B builds

MAX_BN = 256
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID, ACT_TANH, ACT_GELU = range(6)

_VP, _I, _F, _LL, _SZ, _CP = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_size_t,
                              ctypes.c_char_p)
# name -> (restype, argtypes); must list every symbol declared in include/propainter_b200.h
_SIGNATURES = {
    "pp_last_error": (_CP, []),
    "pp_version": (_CP, []),
    "pp_create": (_I, [_I, _VP, _SZ, ctypes.POINTER(_VP)]),
    "pp_destroy": (_I, [_VP]),
    "pp_set_workspace": (_I, [_VP, _VP, _SZ]),
    "pp_comm_unique_id": (_I, [_VP]),
    "pp_comm_init": (_I, [_VP, _VP, _I, _I]),
    "pp_comm_destroy": (_I, [_VP]),
    "pp_comm_all_gather_rows": (_I, [_VP, _VP, ctypes.POINTER(_LL), _SZ, _I, _I, _VP]),
    "pp_register_conv": (_I, [_VP, _CP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I]),
    "pp_register_tensor": (_I, [_VP, _CP, _VP, _SZ]),
    "pp_set_conv_macs": (_I, [_VP, _CP, ctypes.c_double]),
    "pp_raft_bidir": (_I, [_VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    "pp_flow_complete": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP, _VP, _VP]),
    "pp_flow_complete_dist": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP, _VP, _I, _I, _VP]),
    "pp_image_propagate": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP, _VP, _VP]),
    "pp_gen_begin": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "pp_gen_begin_subset": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, ctypes.c_char_p, _VP]),
    "pp_gen_window": (_I, [_VP, ctypes.POINTER(_I), _I, _I, _VP, _VP]),
    "pp_gen_run": (_I, [_VP, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I), _I, _VP, _VP]),
    "pp_gen_end": (_I, [_VP]),
    "pp_composite": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "pp_preprocess": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "pp_preprocess_resize": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "pp_preprocess_u8": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "pp_host_quantize_u8": (_I, [_VP, _VP, _LL, _I]),
    "pp_postprocess": (_I, [_VP, _VP, _VP, _LL, _VP]),
    "pp_launch_count": (_LL, [_VP]),
    "pp_workspace_peak": (_SZ, [_VP]),
    "pp_profile_enable": (_I, [_VP, _I]),
    "pp_profile_dump": (_I, [_VP, ctypes.c_char_p, _SZ]),
    "pp_op_conv": (_I, [_VP, _CP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _F, _VP, _VP, _VP]),
    "pp_op_corr_lookup": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _LL, _I, _I, _VP]),
    "pp_op_imgprop_step": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP]),
    "pp_op_dcn_sample": (_I, [_VP, _VP, _VP, _VP, _F, _VP, _I, _I, _I, _I, _I, _VP]),
    "pp_op_attention": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
}


def load_library() -> ctypes.CDLL:
    """Load the CUDA library (built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_PKG_DIR, 'csrc')}` "
                "(there is no CPU or PyTorch fallback for the ProPainter hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def exported_symbols():
    return sorted(_SIGNATURES)


# ----------------------------------------------------------------------------------------------
# weight packing (host side)
# ----------------------------------------------------------------------------------------------


def choose_bn(cout: int) -> Tuple[int, int]:
    """N-tile of the tcgen05 GEMM: tiles of <= MAX_BN columns (multiple of 16) with the least padding.

    MAX_BN = 256: two 256-column fp32 accumulators fill the 512 TMEM columns of the persistent CTA, and a wide
    N tile halves the im2col (A operand) traffic per output for the Cout >= 256 layers."""
    best = None
    t0 = (cout + MAX_BN - 1) // MAX_BN
    for n_tiles in (t0, t0 + 1):
        bn = ((cout + n_tiles - 1) // n_tiles + 15) // 16 * 16
        if bn > MAX_BN:
            continue
        cand = (bn * n_tiles, n_tiles, bn)
        if best is None or cand < best:
            best = cand
    return best[2], best[0]


def pack_conv_weight(w: torch.Tensor, groups: int = 1, cin_map=None):
    """[Cout, Cin_g, kh, kw] fp32 -> swizzled B-operand image + metadata.

    K is ordered (ky, kx, ci) with ci running over the *kernel's* input channels: ``cin_map[ci]`` is
    the reference input channel feeding kernel channel ci, or -1 for a zero (padding) channel.
    Layout: [groups][K_pad/64][cout_g_pad] rows of 64 fp16; inside each 128-byte row the 16-byte chunk
    c is stored at position c ^ (row & 7) (the 128B swizzle the UMMA descriptor expects)."""
    w = w.detach().float().cpu()
    cout, cin_ref, kh, kw = w.shape
    if cin_map is None:
        cin_map = list(range(cin_ref)) + [-1] * ((-cin_ref) % 8)
    assert len(cin_map) % 8 == 0
    cin_k = len(cin_map)
    idx = torch.tensor([max(i, 0) for i in cin_map], dtype=torch.long)
    keep = torch.tensor([1.0 if i >= 0 else 0.0 for i in cin_map])
    wk = w[:, idx] * keep.view(1, -1, 1, 1)                # [Cout, cin_k, kh, kw]
    wk = wk.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin_k)   # K = (ky, kx, ci)
    cout_g = cout // groups
    bn, cout_g_pad = choose_bn(cout_g)
    K = wk.shape[1]
    K_pad = (K + 63) // 64 * 64
    buf = torch.zeros(groups, cout_g_pad, K_pad)
    buf[:, :cout_g, :K] = wk.view(groups, cout_g, K)
    num_kc = K_pad // 64
    buf = buf.view(groups, cout_g_pad, num_kc, 8, 8).permute(0, 2, 1, 3, 4).contiguous()  # [G, kc, row, chunk, 8]
    rows = torch.arange(cout_g_pad)
    pos = torch.arange(8).view(1, 8) ^ (rows.view(-1, 1) & 7)                               # position p holds chunk p^(r&7)
    buf = torch.gather(buf, 3, pos.view(1, 1, cout_g_pad, 8, 1).expand(groups, num_kc, cout_g_pad, 8, 8))
    meta = dict(cout_g=cout_g, cout_g_pad=cout_g_pad, bn=bn, cin_g=cin_k, kh=kh, kw=kw, groups=groups)
    return buf.to(torch.float16).contiguous(), meta


def _fold_bn(w, b, sd, p, eps=1e-5):
    scale = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + eps)
    return w * scale.view(-1, 1, 1, 1), (b - sd[p + ".running_mean"]) * scale + sd[p + ".bias"]


def _pad_map(n_real: int, total: int):
    return list(range(n_real)) + [-1] * (total - n_real)


def build_layers(raft_sd, rfc_sd, gen_sd):
    """-> (convs: name -> (weight[Cout,Cin,kh,kw], bias, groups, cin_map), tensors: name -> fp32 tensor)."""
    convs: Dict[str, tuple] = {}
    tens: Dict[str, torch.Tensor] = {}

    def add(name, w, b, groups=1, cin_map=None, macs=None):
        # macs: multiply-adds per output pixel of the reference layer (default: the weight tensor as given)
        convs[name] = (w.float(), None if b is None else b.float(), groups, cin_map,
                       float(w.shape[0] * w.shape[1] * w.shape[2] * w.shape[3]) if macs is None else float(macs))

    # ------------------------------------------------------------------ RAFT
    r = {(k[7:] if k.startswith("module.") else k): v.float() for k, v in raft_sd.items()}
    Wspec.check_state_dict(r, Wspec.raft_spec())
    for net, bn in (("fnet", False), ("cnet", True)):
        def cv(dst, src, norm=None, cin_map=None):
            w, b = r[f"{net}.{src}.weight"], r[f"{net}.{src}.bias"]
            if bn and norm is not None:
                w, b = _fold_bn(w, b, r, f"{net}.{norm}")
            add(f"raft.{net}.{dst}", w, b, 1, cin_map)
        cv("conv1", "conv1", "norm1", _pad_map(3, 8))
        for li in (1, 2, 3):
            for bi in (0, 1):
                q = f"layer{li}.{bi}."
                cv(q + "conv1", q + "conv1", q + "norm1")
                cv(q + "conv2", q + "conv2", q + "norm2")
                if li > 1 and bi == 0:
                    cv(q + "downsample", q + "downsample.0", q + "norm3")
        cv("conv2", "conv2")
    u = "update_block."
    add("raft.update.convc1", r[u + "encoder.convc1.weight"], r[u + "encoder.convc1.bias"], 1, _pad_map(324, 328))
    add("raft.update.convc2", r[u + "encoder.convc2.weight"], r[u + "encoder.convc2.bias"])
    # convf1 (7x7 over the 2-channel flow) as a linear layer over explicit 7x7x2 patches in (ky, kx, channel) order,
    # zero-padded 98 -> 128 (kernels_raft.cu: flow_patch7x7)
    wf1 = r[u + "encoder.convf1.weight"]                                   # [128, 2, 7, 7]
    wf1 = torch.cat([wf1.permute(0, 2, 3, 1).reshape(128, 98), torch.zeros(128, 30)], 1).view(128, 128, 1, 1)
    add("raft.update.convf1", wf1, r[u + "encoder.convf1.bias"], 1, None, macs=128 * 98)
    add("raft.update.convf2", r[u + "encoder.convf2.weight"], r[u + "encoder.convf2.bias"])
    add("raft.update.conv", r[u + "encoder.conv.weight"], r[u + "encoder.conv.bias"])
    for s in ("1", "2"):
        add("raft.update.gru.zr" + s, torch.cat([r[u + f"gru.convz{s}.weight"], r[u + f"gru.convr{s}.weight"]], 0),
            torch.cat([r[u + f"gru.convz{s}.bias"], r[u + f"gru.convr{s}.bias"]], 0))
        add("raft.update.gru.q" + s, r[u + f"gru.convq{s}.weight"], r[u + f"gru.convq{s}.bias"])
    add("raft.update.fh1", r[u + "flow_head.conv1.weight"], r[u + "flow_head.conv1.bias"])
    add("raft.update.fh2", r[u + "flow_head.conv2.weight"], r[u + "flow_head.conv2.bias"])
    add("raft.update.mask0", r[u + "mask.0.weight"], r[u + "mask.0.bias"])
    add("raft.update.mask2", r[u + "mask.2.weight"], r[u + "mask.2.bias"])

    # ------------------------------------------------------------------ flow completion
    f = {k: v.float() for k, v in rfc_sd.items()}
    Wspec.check_state_dict(f, Wspec.rfc_spec())
    add("rfc.downsample", f["downsample.0.weight"][:, :, 0], f["downsample.0.bias"], 1, _pad_map(3, 8))
    for enc in ("encoder1", "encoder2"):
        for i in (0, 2):
            add(f"rfc.{enc}.{i}.conv1", f[f"{enc}.{i}.conv1.0.weight"][:, :, 0], f[f"{enc}.{i}.conv1.0.bias"])
            add(f"rfc.{enc}.{i}.conv2", f[f"{enc}.{i}.conv2.0.weight"][:, :, :, :, 0], f[f"{enc}.{i}.conv2.0.bias"])
    for j, i in enumerate((0, 2, 4)):
        add(f"rfc.mid.{j}", f[f"mid_dilation.{i}.weight"][:, :, 0], f[f"mid_dilation.{i}.bias"])

    def add_align(dst, sd, src):
        for j, i in enumerate((0, 2, 4, 6)):
            w = sd[f"{src}.conv_offset.{i}.weight"]
            cm = None if w.shape[1] % 8 == 0 else _pad_map(w.shape[1], (w.shape[1] + 7) // 8 * 8)
            add(f"{dst}.offset.{j}", w, sd[f"{src}.conv_offset.{i}.bias"], 1, cm)
        w = sd[src + ".weight"]                                        # [Cout, Cin, 3, 3] -> K = (tap, ci)
        add(f"{dst}.dcn", w.permute(0, 2, 3, 1).reshape(w.shape[0], -1, 1, 1), sd[src + ".bias"])

    fp = "feat_prop_module."
    for d in ("backward_", "forward_"):
        add_align(f"rfc.fp.{d}", f, fp + "deform_align." + d)
        add(f"rfc.fp.{d}.backbone.0", f[fp + f"backbone.{d}.0.weight"], f[fp + f"backbone.{d}.0.bias"])
        add(f"rfc.fp.{d}.backbone.1", f[fp + f"backbone.{d}.2.weight"], f[fp + f"backbone.{d}.2.bias"])
    add("rfc.fp.fusion", f[fp + "fusion.weight"], f[fp + "fusion.bias"])
    for dst, src in (("decoder2.0", "decoder2.0"), ("decoder2.deconv", "decoder2.2.conv"), ("decoder1.0", "decoder1.0"),
                     ("decoder1.deconv", "decoder1.2.conv"), ("upsample.0", "upsample.0"),
                     ("upsample.deconv", "upsample.2.conv")):
        add("rfc." + dst, f[src + ".weight"], f[src + ".bias"])

    # ------------------------------------------------------------------ generator
    g = {k: v.float() if v.is_floating_point() else v for k, v in gen_sd.items()}
    Wspec.check_state_dict(g, Wspec.generator_spec())
    enc_groups = {0: 1, 2: 1, 4: 1, 6: 1, 8: 1, 10: 2, 12: 4, 14: 8, 16: 1}
    for i, gr in enc_groups.items():
        w = g[f"encoder.layers.{i}.weight"]
        if i == 14:
            # groups of 80 input / 32 output channels are too small for 64-wide K chunks and 128-column tiles: run the
            # layer dense with block-diagonal weights (8x the MACs, all of them on the TMA halo-tile kernel at >10x the
            # rate).  Kernel channel order = cat(x0[256], previous output[384]); group k owns x0[32k:32k+32] and
            # prev[48k:48k+48] (propainter.py:268-273).
            cout, cg = w.shape[0], w.shape[1]
            nx, npv = 256 // gr, 384 // gr
            assert cg == nx + npv and cout % gr == 0
            dense = torch.zeros(cout, 640, w.shape[2], w.shape[3])
            for k in range(gr):
                rows = slice(k * (cout // gr), (k + 1) * (cout // gr))
                dense[rows, k * nx:(k + 1) * nx] = w[rows, :nx]
                dense[rows, 256 + k * npv:256 + (k + 1) * npv] = w[rows, nx:]
            add("gen.encoder.14", dense, g["encoder.layers.14.bias"], 1, None, macs=w.numel())
            continue
        add(f"gen.encoder.{i}", w, g[f"encoder.layers.{i}.bias"], gr, _pad_map(5, 8) if i == 0 else None)
    for dst, src in (("0", "0.conv"), ("2", "2"), ("4", "4.conv"), ("6", "6")):
        add("gen.decoder." + dst, g[f"decoder.{src}.weight"], g[f"decoder.{src}.bias"])
    add("gen.ss", g["ss.embedding.weight"].view(512, 128, 7, 7), g["ss.embedding.bias"])
    # SoftComp Linear: output column c*49+k -> k*128+c, so the fold kernel reads contiguous channels
    perm = (torch.arange(49).view(49, 1) + 49 * torch.arange(128).view(1, 128)).reshape(-1)
    add("gen.sc.embedding", g["sc.embedding.weight"][perm].view(6272, 512, 1, 1), g["sc.embedding.bias"][perm])
    add("gen.sc.bias_conv", g["sc.bias_conv.weight"], g["sc.bias_conv.bias"])
    for d in ("backward_1", "forward_1"):
        add_align(f"gen.fp.{d}", g, fp + "deform_align." + d)
        add(f"gen.fp.{d}.backbone.0", g[fp + f"backbone.{d}.0.weight"], g[fp + f"backbone.{d}.0.bias"], 1,
            _pad_map(258, 264))
        add(f"gen.fp.{d}.backbone.1", g[fp + f"backbone.{d}.2.weight"], g[fp + f"backbone.{d}.2.bias"])
    add("gen.fp.fuse.0", g[fp + "fuse.0.weight"], g[fp + "fuse.0.bias"], 1, _pad_map(258, 264))
    add("gen.fp.fuse.1", g[fp + "fuse.2.weight"], g[fp + "fuse.2.bias"])
    perm40 = (torch.arange(49).view(49, 1) + 49 * torch.arange(40).view(1, 40)).reshape(-1)
    for b in range(Wspec.N_TRANSFORMER_BLOCKS):
        t = f"transformers.transformer.{b}."
        a = t + "attention."
        o = f"gen.tf.{b}."
        qkv_w = torch.cat([g[a + "query.weight"], g[a + "key.weight"], g[a + "value.weight"]], 0)
        qkv_b = torch.cat([g[a + "query.bias"], g[a + "key.bias"], g[a + "value.bias"]], 0)
        add(o + "qkv", qkv_w.view(1536, 512, 1, 1), qkv_b)
        add(o + "kv", qkv_w[512:].reshape(1024, 512, 1, 1), qkv_b[512:])
        add(o + "proj", g[a + "proj.weight"].view(512, 512, 1, 1), g[a + "proj.bias"])
        add(o + "fc1", g[t + "mlp.fc1.0.weight"][perm40].view(1960, 512, 1, 1), g[t + "mlp.fc1.0.bias"][perm40])
        add(o + "fc2", g[t + "mlp.fc2.1.weight"].view(512, 40, 7, 7), g[t + "mlp.fc2.1.bias"])
        for n in ("norm1", "norm2"):
            tens[o + n + ".weight"] = g[t + n + ".weight"].float()
            tens[o + n + ".bias"] = g[t + n + ".bias"].float()
        tens[o + "pool.weight"] = g[a + "pool_layer.weight"].reshape(512, 16).float().t().contiguous()   # [tap][C]
        tens[o + "pool.bias"] = g[a + "pool_layer.bias"].float()
        expect = torch.from_numpy(Wspec.rolled_valid_indices())
        if not torch.equal(g[a + "valid_ind_rolled"].cpu().long(), expect):
            raise ValueError("checkpoint's valid_ind_rolled differs from the 5x9 window ring this engine implements")
    return convs, tens


# ----------------------------------------------------------------------------------------------
# engine
# ----------------------------------------------------------------------------------------------


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Engine:
    """One engine per process/GPU.  Owns the workspace arena and the packed weights."""

    MIN_WORKSPACE = 256 << 20

    def __init__(self, device: torch.device | str | int = "cuda:0", workspace_gb: float | None = None):
        """``workspace_gb``: fixed size of the scratch arena; None = start at 256 MiB and let ``reserve_for_clip``
        size it from (T, H, W) before each clip (what the node path does)."""
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("the ProPainter B200 engine needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._keep = []
        self._total_mem = torch.cuda.get_device_properties(self.device).total_memory
        self.fixed_workspace = workspace_gb is not None
        nbytes = self.MIN_WORKSPACE if workspace_gb is None else int(min(workspace_gb * (1 << 30), self._total_mem * 0.8))
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        h = ctypes.c_void_p()
        self._check(self.lib.pp_create(self.device.index, _ptr(self.workspace), nbytes, ctypes.byref(h)))
        self.h = h
        self.conv_meta: Dict[str, dict] = {}

    # -- workspace sizing
    @staticmethod
    def clip_workspace_bytes(T: int, H: int, W: int) -> int:
        """Arena size that lets every stage of a T x H x W clip run in its widest batching (measured peaks: 24.3 GB at
        80 x 640x360, 92 GB at 80 x 1280x720 => ~1.3 kB per frame-pixel; clips beyond ~100 frames are processed in
        sub-batches of windows / chunks of sub-videos, so the estimate saturates there)."""
        return int(1400 * min(T, 100) * H * W + (2 << 30))

    def set_workspace_bytes(self, nbytes: int) -> None:
        """Re-allocate the arena (no generator session may be open)."""
        nbytes = max(int(nbytes), self.MIN_WORKSPACE)
        torch.cuda.current_stream(self.device).synchronize()
        self.workspace = None                       # release before allocating: never hold old + new together
        torch.cuda.empty_cache()
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._check(self.lib.pp_set_workspace(self.h, _ptr(ws), nbytes))
        self.workspace = ws

    def reserve_for_clip(self, T: int, H: int, W: int) -> int:
        """Grow the arena for a clip when it was created without a fixed size.  Capped at 80 % of the device memory;
        when the cap (or a fixed size) is below the estimate the stages fall back to smaller batches
        (RAFT pair batches, encoder / decoder frame chunks, sub-batches of sliding windows)."""
        if self.fixed_workspace:
            return self.workspace.numel()
        free, _ = torch.cuda.mem_get_info(self.device)
        have = self.workspace.numel()
        want = min(self.clip_workspace_bytes(T, H, W), int(self._total_mem * 0.8), int((free + have) * 0.9))
        if want > have:
            self.set_workspace_bytes(want)
        return self.workspace.numel()

    def release_workspace(self) -> None:
        """Shrink the arena back to its minimum (gives the HBM back between node executions when asked to)."""
        if not self.fixed_workspace and self.workspace.numel() > self.MIN_WORKSPACE:
            self.set_workspace_bytes(self.MIN_WORKSPACE)

    # -- helpers
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError("propainter_b200: " + self.lib.pp_last_error().decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pp_destroy(self.h)
            self.h = None
        self.workspace = None
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights
    def register_conv(self, name, w, b, groups=1, cin_map=None, macs=None):
        packed, meta = pack_conv_weight(w, groups, cin_map)
        packed = packed.to(self.device)
        bias = None if b is None else b.detach().float().contiguous().to(self.device)
        self._keep += [packed, bias]
        self.conv_meta[name] = meta
        self._check(self.lib.pp_register_conv(self.h, name.encode(), _ptr(packed), _ptr(bias), meta["cout_g"],
                                              meta["cout_g_pad"], meta["bn"], meta["cin_g"], meta["kh"], meta["kw"],
                                              meta["groups"]))
        self._check(self.lib.pp_set_conv_macs(self.h, name.encode(), float(w.numel() if macs is None else macs)))

    def register_tensor(self, name, t):
        t = t.detach().float().contiguous().to(self.device)
        self._keep.append(t)
        self._check(self.lib.pp_register_tensor(self.h, name.encode(), _ptr(t), t.numel() * 4))

    # Stride-1 k>1 layers whose input channels are not a multiple of 64: kernel channels zero-padded to the next
    # multiple so they run on the TMA halo-tile kernel (64-channel K chunks); the activation tensors keep their real
    # channel count, TMA zero-fills the rest of the last segment (PPConvSeg.cvalid)
    PAD64_CONVS = ("rfc.encoder1.0.conv1", "rfc.encoder1.0.conv2", "rfc.upsample.0", "rfc.upsample.deconv") + tuple(
        f"raft.{net}.layer2.{blk}" for net in ("fnet", "cnet") for blk in ("0.conv2", "1.conv1", "1.conv2")) + (  # 96 ch
        # cat(features[256], mask/flow[8]) = 264 kernel channels -> 320: the 8-channel tail segment fills one chunk
        "gen.fp.backward_1.offset.0", "gen.fp.forward_1.offset.0", "gen.fp.backward_1.backbone.0",
        "gen.fp.forward_1.backbone.0", "gen.fp.fuse.0")

    def load_weights(self, raft_sd, rfc_sd, gen_sd):
        convs, tens = build_layers(raft_sd, rfc_sd, gen_sd)
        for name, (w, b, groups, cin_map, macs) in convs.items():
            if name in self.PAD64_CONVS:
                cin_map = list(cin_map) if cin_map is not None else list(range(w.shape[1]))
                cin_map += [-1] * ((-len(cin_map)) % 64)
            self.register_conv(name, w, b, groups, cin_map, macs)
        for name, t in tens.items():
            self.register_tensor(name, t)
        return self

    # -- stages (float32 contiguous CUDA tensors in the reference's layouts)
    def _f32(self, t):
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def raft_bidir(self, frames: torch.Tensor, iters: int, out=None):
        """frames [T,3,H,W] in [-1,1] -> (flows_f, flows_b) [T-1,2,H,W] (written into `out` when given: contiguous
        float32 views, e.g. a rank's shard of the full flow buffers)."""
        frames = self._f32(frames)
        T, _, H, W = frames.shape
        if out is not None:
            ff, fb = out
            assert ff.is_contiguous() and fb.is_contiguous() and ff.dtype == torch.float32 and ff.shape == (T - 1, 2, H, W)
        else:
            ff = torch.empty(T - 1, 2, H, W, device=self.device, dtype=torch.float32)
            fb = torch.empty_like(ff)
        self._check(self.lib.pp_raft_bidir(self.h, _ptr(frames), T, H, W, int(iters), _ptr(ff), _ptr(fb), self._stream()))
        return ff, fb

    def flow_complete(self, flows_f, flows_b, flow_masks):
        flows_f, flows_b, flow_masks = self._f32(flows_f), self._f32(flows_b), self._f32(flow_masks)
        T, _, H, W = flow_masks.shape
        assert flows_f.shape[0] == T - 1
        of, ob = torch.empty_like(flows_f), torch.empty_like(flows_b)
        self._check(self.lib.pp_flow_complete(self.h, _ptr(flows_f), _ptr(flows_b), _ptr(flow_masks), T, H, W, _ptr(of),
                                              _ptr(ob), self._stream()))
        return of, ob

    def flow_complete_dist(self, flows_f, flows_b, flow_masks, team_first: int, team_size: int, out=None):
        """Collective flow completion of one chunk by the ranks [team_first, team_first + team_size) (see
        pp_flow_complete_dist); every team member gets the full completed flows."""
        flows_f, flows_b, flow_masks = self._f32(flows_f), self._f32(flows_b), self._f32(flow_masks)
        T, _, H, W = flow_masks.shape
        assert flows_f.shape[0] == T - 1
        of, ob = out if out is not None else (torch.empty_like(flows_f), torch.empty_like(flows_b))
        assert of.is_contiguous() and ob.is_contiguous() and of.dtype == torch.float32
        self._check(self.lib.pp_flow_complete_dist(self.h, _ptr(flows_f), _ptr(flows_b), _ptr(flow_masks), T, H, W,
                                                   _ptr(of), _ptr(ob), int(team_first), int(team_size), self._stream()))
        return of, ob

    def image_propagate(self, frames, masks, flows_f, flows_b):
        frames, masks, flows_f, flows_b = map(self._f32, (frames, masks, flows_f, flows_b))
        T, _, H, W = frames.shape
        uf, um = torch.empty_like(frames), torch.empty_like(masks)
        self._check(self.lib.pp_image_propagate(self.h, _ptr(frames), _ptr(masks), _ptr(flows_f), _ptr(flows_b), T, H, W,
                                                _ptr(uf), _ptr(um), self._stream()))
        return uf, um

    def gen_begin(self, updated_frames, masks_dilated, updated_masks, flows_f, flows_b, frames_needed=None):
        """Open a generator session over the clip.  ``frames_needed``: frame ids to encode (default all); the windows
        given to gen_run must only touch those."""
        a = [self._f32(x) for x in (updated_frames, masks_dilated, updated_masks, flows_f, flows_b)]
        T, _, H, W = a[0].shape
        self._gen_shape = (T, H, W)
        self._gen_inputs = a  # keep alive for the session
        self._gen_needed = None if frames_needed is None else set(int(i) for i in frames_needed)
        if frames_needed is None:
            self._check(self.lib.pp_gen_begin(self.h, *[_ptr(x) for x in a], T, H, W, self._stream()))
        else:
            need = bytes(1 if i in self._gen_needed else 0 for i in range(T))
            self._check(self.lib.pp_gen_begin_subset(self.h, *[_ptr(x) for x in a], T, H, W, need, self._stream()))

    def gen_window(self, frame_ids, l_t: int) -> torch.Tensor:
        """-> fp16 [l_t,H,W,4] (rgb in [-1,1], lane 3 unused)."""
        T, H, W = self._gen_shape
        ids = (ctypes.c_int * len(frame_ids))(*[int(i) for i in frame_ids])
        pred = torch.empty(l_t, H, W, 4, device=self.device, dtype=torch.float16)
        self._check(self.lib.pp_gen_window(self.h, ids, len(frame_ids), int(l_t), _ptr(pred), self._stream()))
        return pred

    @staticmethod
    def gen_slot_bytes(H: int, W: int) -> int:
        """Workspace one (window, frame) slot of pp_gen_run needs (generator.cu): window-major features and the token /
        qkv / FFN rows of the transformer (all slots alive together), plus an upper bound of the feature-propagation
        buffers (4 feature-sized tensors per LOCAL slot of the largest equal-length group, and the per-step condition /
        offset / sampled-column tensors, one set per window ~ 1/8 of a slot)."""
        p4 = (H // 4) * (W // 4)
        gh, gw = (H // 4 + 6 - 7) // 3 + 1, (W // 4 + 6 - 7) // 3 + 1
        rows_pad = -(-gh // 5) * 5 * (-(-gw // 9) * 9)
        xfmr = p4 * (256 + 80) + gh * gw * 2 * (512 * 3 + 1960) + rows_pad * 2 * (512 + 1536)
        featprop = p4 * 2 * (128 * 4) + p4 * 2 * (264 + 432 + 1152 + 128 * 3) // 8
        return int(1.25 * (xfmr + featprop))

    def gen_batches(self, windows, budget_bytes: int, shape=None):
        """Split the schedule into consecutive sub-batches whose slots fit `budget_bytes` (windows are independent;
        the composite order is the window order, which consecutive sub-batches keep)."""
        T, H, W = shape if shape is not None else self._gen_shape
        per_slot = self.gen_slot_bytes(H, W)
        out, cur, used = [], [], 0
        for w in windows:
            need = (len(w[0]) + len(w[1])) * per_slot
            if cur and used + need > budget_bytes:
                out.append(cur)
                cur, used = [], 0
            cur.append(w)
            used += need
        if cur:
            out.append(cur)
        return out

    def gen_run(self, windows) -> torch.Tensor:
        """Sliding windows in batched passes.  windows = [(neighbor_ids, ref_ids), ...]
        -> fp16 [sum(len(neighbor_ids)), H, W, 4] in window order.

        One engine pass covers as many windows as the workspace holds (all 16 of an 80-frame 640x360 clip); a long
        or large clip is split into consecutive sub-batches, down to one window per pass, before giving up --
        the reference runs one window at a time, so anything it can process this can too."""
        T, H, W = self._gen_shape
        enc_bytes = T * (H // 4) * (W // 4) * (256 + 32) + (64 << 20)          # the session's resident part
        decoder_reserve = 24 * H * W * 2 * 200                                   # room for a useful decoder chunk
        budget = max(self.workspace.numel() - enc_bytes - decoder_reserve, self.gen_slot_bytes(H, W))
        out = [self._gen_run_or_split(b) for b in self.gen_batches(list(windows), budget)]
        return out[0] if len(out) == 1 else torch.cat(out, 0)

    def _gen_run_or_split(self, windows) -> torch.Tensor:
        try:
            return self._gen_run_once(windows)
        except RuntimeError as ex:
            if "workspace" not in str(ex) or len(windows) == 1:
                raise
        half = len(windows) // 2                     # the estimate was too optimistic: halve and retry
        return torch.cat([self._gen_run_or_split(windows[:half]), self._gen_run_or_split(windows[half:])], 0)

    gen_run_calls = 0       # engine passes issued by gen_run (1 per clip unless the workspace forced sub-batches)

    def _gen_run_once(self, windows) -> torch.Tensor:
        self.gen_run_calls += 1
        T, H, W = self._gen_shape
        flat, wt, wl = [], [], []
        if getattr(self, "_gen_needed", None) is not None:
            missing = {int(i) for nb, refs in windows for i in list(nb) + list(refs)} - self._gen_needed
            if missing:
                raise ValueError(f"gen_run: frames {sorted(missing)} were not encoded by gen_begin(frames_needed=...)")
        for nb, refs in windows:
            flat += [int(i) for i in nb] + [int(i) for i in refs]
            wt.append(len(nb) + len(refs))
            wl.append(len(nb))
        arr = lambda v: (ctypes.c_int * len(v))(*v)
        pred = torch.empty(sum(wl), H, W, 4, device=self.device, dtype=torch.float16)
        self._check(self.lib.pp_gen_run(self.h, arr(flat), arr(wt), arr(wl), len(windows), _ptr(pred), self._stream()))
        return pred

    def gen_end(self):
        self._check(self.lib.pp_gen_end(self.h))
        self._gen_inputs = None

    def composite(self, pred, masks_dilated, orig_u8, comp_u8, frame_ids_dev, first_visit_dev, half_math=False):
        """half_math: reproduce the half-precision roundings of the reference's fp16="enable" mode before the uint8
        truncation (reference propainter_inference.py:285-286 on a half tensor)."""
        l_t, H, W, _ = pred.shape
        self._check(self.lib.pp_composite(self.h, _ptr(pred), _ptr(masks_dilated), _ptr(orig_u8), _ptr(comp_u8),
                                          _ptr(frame_ids_dev), _ptr(first_visit_dev), l_t, H, W, int(bool(half_math)),
                                          self._stream()))

    def preprocess(self, image, mask, flow_mask_dilates: int, mask_dilates: int, process_size=None):
        """Device version of convert_image_to_frames + prepare_frames_and_masks (reference utils/image_utils.py:98-197).
        image [T,H,W,3] float 0..1, mask [T or 1,H,W] float32 (host or device); ``process_size`` = (width, height) to
        resize to (PIL's 8-bit bicubic resampler, reproduced bit for bit on the device), default: the input size.
        -> frames [1,T,3,h,w], flow_masks [1,T,1,h,w], masks_dilated [1,T,1,h,w] (float32), originals uint8 [T,h,w,3]."""
        T, H, W, _ = image.shape
        ow, oh = (W, H) if process_size is None else (int(process_size[0]), int(process_size[1]))
        if image.device.type == "cpu" and mask.device.type == "cpu" and image.dtype == torch.float32 and mask.dtype == torch.float32:
            # host tensors (what ComfyUI hands a node): the float -> uint8 truncation is the first thing the reference does
            # with them, so do it on the host cores straight into page-locked staging buffers and move 1/4 of the bytes
            return self._preprocess_host(image.contiguous(), mask.contiguous(), flow_mask_dilates, mask_dilates, ow, oh)
        img = image.to(self.device, torch.float32, non_blocking=True).contiguous()
        msk = mask.to(self.device, torch.float32, non_blocking=True).contiguous()
        orig = torch.empty(T, oh, ow, 3, device=self.device, dtype=torch.uint8)
        frames = torch.empty(T, 3, oh, ow, device=self.device, dtype=torch.float32)
        fm = torch.empty(T, 1, oh, ow, device=self.device, dtype=torch.float32)
        md = torch.empty_like(fm)
        if (ow, oh) == (W, H):
            self._check(self.lib.pp_preprocess(self.h, _ptr(img), _ptr(msk), msk.shape[0], T, H, W, int(flow_mask_dilates),
                                               int(mask_dilates), _ptr(orig), _ptr(frames), _ptr(fm), _ptr(md),
                                               self._stream()))
        else:
            self._check(self.lib.pp_preprocess_resize(self.h, _ptr(img), _ptr(msk), msk.shape[0], T, H, W, oh, ow,
                                                      int(flow_mask_dilates), int(mask_dilates), _ptr(orig), _ptr(frames),
                                                      _ptr(fm), _ptr(md), self._stream()))
        return frames.unsqueeze(0), fm.unsqueeze(0), md.unsqueeze(0), orig

    def _preprocess_host(self, image, mask, flow_mask_dilates, mask_dilates, ow, oh):
        T, H, W, _ = image.shape
        threads = min(os.cpu_count() or 1, int(os.environ.get("PP_HOST_THREADS", 16)))
        try:
            img8 = torch.empty(image.shape, dtype=torch.uint8, pin_memory=True)
            msk8 = torch.empty(mask.shape, dtype=torch.uint8, pin_memory=True)
        except RuntimeError:        # locked-memory limit: pageable staging still moves 1/4 of the bytes
            img8 = torch.empty(image.shape, dtype=torch.uint8)
            msk8 = torch.empty(mask.shape, dtype=torch.uint8)
        self._check(self.lib.pp_host_quantize_u8(_ptr(image), _ptr(img8), image.numel(), threads))
        img8d = img8.to(self.device, non_blocking=True)
        self._check(self.lib.pp_host_quantize_u8(_ptr(mask), _ptr(msk8), mask.numel(), threads))
        msk8d = msk8.to(self.device, non_blocking=True)
        orig = torch.empty(T, oh, ow, 3, device=self.device, dtype=torch.uint8)
        frames = torch.empty(T, 3, oh, ow, device=self.device, dtype=torch.float32)
        fm = torch.empty(T, 1, oh, ow, device=self.device, dtype=torch.float32)
        md = torch.empty_like(fm)
        self._check(self.lib.pp_preprocess_u8(self.h, _ptr(img8d), _ptr(msk8d), mask.shape[0], T, H, W, oh, ow,
                                              int(flow_mask_dilates), int(mask_dilates), _ptr(orig), _ptr(frames), _ptr(fm),
                                              _ptr(md), self._stream()))
        self._keep_staging = (img8, msk8)      # the async copies read them; released at the next call
        return frames.unsqueeze(0), fm.unsqueeze(0), md.unsqueeze(0), orig

    def postprocess(self, comp_u8: torch.Tensor) -> torch.Tensor:
        """uint8 [T,H,W,3] -> float32/255 on the device (handle_output)."""
        out = torch.empty(comp_u8.shape, device=self.device, dtype=torch.float32)
        self._check(self.lib.pp_postprocess(self.h, _ptr(comp_u8), _ptr(out), comp_u8.numel(), self._stream()))
        return out

    # -- multi-GPU exchange (NCCL communicator inside the C library)
    rank, world = 0, 1

    def comm_unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        self._check(self.lib.pp_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == 128
        self._check(self.lib.pp_comm_init(self.h, ctypes.create_string_buffer(unique_id, 128), int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def comm_destroy(self):
        self._check(self.lib.pp_comm_destroy(self.h))
        self.rank, self.world = 0, 1

    def comm_all_gather_rows(self, buf: torch.Tensor, rows, first_rank: int = 0):
        """In-place all-gather along dim 0 of the contiguous tensor `buf` among ranks
        [first_rank, first_rank + len(rows)): member m owns rows [sum(rows[:m]), +rows[m])."""
        assert buf.is_contiguous() and buf.shape[0] == sum(rows)
        row_bytes = buf[0].numel() * buf.element_size() if buf.shape[0] else 0
        if len(rows) <= 1 or row_bytes == 0:
            return buf
        arr = (ctypes.c_longlong * len(rows))(*[int(r) for r in rows])
        self._check(self.lib.pp_comm_all_gather_rows(self.h, _ptr(buf), arr, row_bytes, int(first_rank), len(rows),
                                                     self._stream()))
        return buf

    @property
    def launch_count(self) -> int:
        return int(self.lib.pp_launch_count(self.h))

    @property
    def workspace_peak(self) -> int:
        return int(self.lib.pp_workspace_peak(self.h))

    def profile_enable(self, on: bool):
        self._check(self.lib.pp_profile_enable(self.h, int(on)))

    def profile_dump(self):
        """-> {kernel name: dict(count, ms, rows, flops, bytes)} since profile_enable(True)."""
        buf = ctypes.create_string_buffer(1 << 20)
        self._check(self.lib.pp_profile_dump(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms, rows, flops, nbytes = line.split("\t")
            out[name] = dict(count=int(n), ms=float(ms), rows=float(rows), flops=float(flops), bytes=float(nbytes))
        return out

    # -- single operators (tests / micro-benchmarks)
    def op_conv(self, name, x_nhwc, stride=1, pad=0, dil=1, replicate=False, act=ACT_NONE, slope=0.0, residual=None):
        m = self.conv_meta[name]
        N, H, W, _ = x_nhwc.shape
        kh, kw = m["kh"], m["kw"]
        OH = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        OW = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        out = torch.empty(N, OH, OW, m["cout_g"] * m["groups"], device=self.device, dtype=torch.float16)
        self._check(self.lib.pp_op_conv(self.h, name.encode(), _ptr(x_nhwc), N, H, W, stride, pad, dil, int(replicate),
                                        act, float(slope), _ptr(residual), _ptr(out), self._stream()))
        return out

    def op_corr_lookup(self, levels, coords, h8, w8):
        nq = coords.shape[0]
        out = torch.empty(nq, 328, device=self.device, dtype=torch.float16)
        self._check(self.lib.pp_op_corr_lookup(self.h, *[_ptr(l) for l in levels], _ptr(coords), _ptr(out), nq, h8, w8,
                                               self._stream()))
        return out

    def op_imgprop_step(self, cur4, prop4, flow_prop, flow_check):
        H, W, _ = cur4.shape
        out = torch.empty_like(cur4)
        self._check(self.lib.pp_op_imgprop_step(self.h, _ptr(cur4), _ptr(prop4), _ptr(out), _ptr(flow_prop),
                                                _ptr(flow_check), H, W, self._stream()))
        return out

    def op_dcn_sample(self, x, offs, flow, max_mag: float, tiled: bool):
        """x [N,H,W,C] fp16, offs [N,H,W,432] fp16, flow [N,H,W,2] fp16 or None -> cols [N*H*W, 9*C] fp16."""
        N, H, W, C = x.shape
        cols = torch.empty(N * H * W, 9 * C, device=self.device, dtype=torch.float16)
        self._check(self.lib.pp_op_dcn_sample(self.h, _ptr(x), _ptr(offs), _ptr(flow), float(max_mag), _ptr(cols), N, H, W, C,
                                              int(tiled), self._stream()))
        return cols

    def op_attention(self, qkv, pkv, win_flags, t, gh, gw, n_pool, parity):
        out = torch.zeros(t, gh, gw, 512, device=self.device, dtype=torch.float16)
        self._check(self.lib.pp_op_attention(self.h, _ptr(qkv), _ptr(pkv), _ptr(out), _ptr(win_flags), t, gh, gw, n_pool,
                                             parity, self._stream()))
        return out
