"""Pin oracle/propainter_oracle.py to outputs of the REAL reference (tests/golden/reference_outputs.npz).

The fixtures were produced by tests/golden/make_golden.py, which imports the unmodified reference in
the build container; inputs are regenerated from seeds here.  CPU only, fp32.
"""
import numpy as np
import torch

from comfyui_propainter_nodes_b200 import weights as Wt
from comfyui_propainter_nodes_b200.utils import image_utils as IU
from oracle import propainter_oracle as O
from tests.golden import cases


def _close(a, ref, atol):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    err = np.abs(a.astype(np.float64) - ref.astype(np.float64)).max()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    assert err <= atol, f"max abs err {err} > {atol}"


def test_raft_matches_reference(golden):
    with torch.no_grad():
        ff, fb = O.raft_bidirectional(Wt.synthetic_raft_state_dict(), cases.raft_case(), cases.RAFT_ITERS)
    _close(ff, golden["raft_ff"], 2e-3)
    _close(fb, golden["raft_fb"], 2e-3)


def test_flow_completion_matches_reference(golden):
    flows, masks = cases.rfc_case()
    with torch.no_grad():
        f, b = O.rfc_bidirectional(Wt.synthetic_rfc_state_dict(), flows, masks)
    _close(f, golden["rfc_f"], 1e-3)
    _close(b, golden["rfc_b"], 1e-3)


def test_image_propagation_matches_reference(golden):
    frames, m, fl = cases.imgprop_case()
    with torch.no_grad():
        uf, um = O.image_propagation(frames, m, fl, 80)
    _close(uf, golden["imgprop_frames"], 1e-6)
    _close(um, golden["imgprop_masks"], 0)


def test_generator_window_matches_reference(golden):
    g = cases.window_case()
    with torch.no_grad():
        pred = O.inpaint_window(Wt.synthetic_generator_state_dict(), g["frames"], g["flows"], g["masks_in"],
                                g["masks_upd"], g["l_t"])
    _close(pred, golden["window_pred"], 1e-3)


def test_end_to_end_matches_reference(golden):
    e = cases.e2e_case()
    cfg = IU.ImageConfig(e["W"], e["H"], 5, 8, (e["W"], e["H"]), e["T"])
    frames_u8 = IU.convert_image_to_frames(e["image"])
    ft, fm, md, orig = IU.prepare_frames_and_masks(frames_u8, e["mask"], cfg, torch.device("cpu"))
    _close(fm, golden["e2e_flow_masks"], 0)
    _close(md, golden["e2e_masks_dilated"], 0)
    comp, st = O.run_pipeline(Wt.synthetic_raft_state_dict(), Wt.synthetic_rfc_state_dict(),
                              Wt.synthetic_generator_state_dict(), ft, fm, md, orig,
                              raft_iter=e["raft_iter"], subvideo_length=e["subvideo_length"],
                              neighbor_length=e["neighbor_length"], ref_stride=e["ref_stride"],
                              return_stages=True)
    _close(st["pred_flows"][0], golden["e2e_pred_flow_f"], 5e-3)
    # nearest-neighbour propagation may flip single pixels when a coordinate lands on .5 +- 1 ulp
    d = np.abs(st["updated_frames"].numpy() - golden["e2e_updated_frames"])
    assert (d > 1e-4).mean() < 1e-3
    out = np.stack(comp).astype(np.int32)
    ref = golden["e2e_frames_u8"].astype(np.int32)
    assert out.shape == ref.shape
    assert (np.abs(out - ref) > 1).mean() < 2e-3, (np.abs(out - ref) > 1).mean()


def test_window_schedule_defaults():
    """80 frames, defaults: 16 windows, sum(t) = 275, sum(l_t) = 170."""
    s = O.window_schedule(80, 10, 10, 80)
    assert len(s) == 16
    assert sum(len(a) + len(b) for a, b in s) == 275
    assert sum(len(a) for a, _ in s) == 170
    # long video: references limited to ref_num=8 within +-40 frames, up to 9 (quirk 17)
    s = O.window_schedule(240, 10, 10, 80)
    assert len(s) == 48 and max(len(b) for _, b in s) <= 9
