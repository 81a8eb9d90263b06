"""Generate the golden fixtures that pin oracle/propainter_oracle.py to the REAL reference.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

It imports the unmodified reference package with a stub ``comfy.model_management``, loads the seeded
synthetic checkpoints from comfyui_propainter_nodes_b200.weights into the reference's own modules
(strict=True) and stores small per-stage input/output tensors as float16/float32 .npz files.
Inputs are regenerated from seeds by the tests; only reference OUTPUTS are stored.
"""
import os
import sys
import types
import argparse
import tempfile

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root")

import numpy as np
import torch

comfy = types.ModuleType("comfy")
mm = types.ModuleType("comfy.model_management")
mm.get_torch_device = lambda: torch.device("cpu")
comfy.model_management = mm
sys.modules["comfy"] = comfy
sys.modules["comfy.model_management"] = mm

import reference  # noqa: E402
from reference.model.modules.flow_comp_raft import RAFT_bi  # noqa: E402
from reference.model.recurrent_flow_completion import RecurrentFlowCompleteNet  # noqa: E402
from reference.model.propainter import InpaintGenerator  # noqa: E402
from reference import propainter_inference as RI  # noqa: E402
from reference.utils import image_utils as RU  # noqa: E402

from comfyui_propainter_nodes_b200 import weights as Wt  # noqa: E402
from tests.golden import cases  # noqa: E402


def build_models():
    tmp = tempfile.mkdtemp()
    rp = os.path.join(tmp, "raft.pth")
    torch.save(Wt.synthetic_raft_state_dict(), rp)
    raft = RAFT_bi(rp, "cpu")
    rfc = RecurrentFlowCompleteNet()
    rfc.load_state_dict(Wt.synthetic_rfc_state_dict(), strict=True)
    rfc.eval()
    gen = InpaintGenerator()
    gen.load_state_dict(Wt.synthetic_generator_state_dict(), strict=True)
    gen.eval()
    return raft, rfc, gen


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    raft, rfc, gen = build_models()
    out = {}
    with torch.no_grad():
        # ---- RAFT
        fr = cases.raft_case()
        ff, fb = raft(fr, iters=cases.RAFT_ITERS)
        out["raft_ff"], out["raft_fb"] = ff, fb
        # ---- flow completion
        flows, masks = cases.rfc_case()
        pred, _ = rfc.forward_bidirect_flow(flows, masks)
        comb = rfc.combine_flow(flows, pred, masks)
        out["rfc_f"], out["rfc_b"] = comb
        # ---- image propagation
        frames, m, fl = cases.imgprop_case()
        cfg = RI.ProPainterConfig(10, 10, 80, 5, "disable", frames.shape[1], torch.device("cpu"),
                                  (frames.shape[-1], frames.shape[-2]))
        uf, um = RI.image_propagation(gen, frames, m, fl, cfg)
        out["imgprop_frames"], out["imgprop_masks"] = uf, um
        # ---- generator window
        g = cases.window_case()
        pred_img = gen(g["frames"], g["flows"], g["masks_in"], g["masks_upd"], g["l_t"])
        out["window_pred"] = pred_img
        # ---- end-to-end through the reference node-level functions
        e = cases.e2e_case()
        icfg = RU.ImageConfig(e["W"], e["H"], 5, 8, (e["W"], e["H"]), e["T"])
        frames_pil = RU.convert_image_to_frames(e["image"])
        ft, fm, md, orig = RU.prepare_frames_and_masks(frames_pil, e["mask"], icfg, torch.device("cpu"))
        out["e2e_flow_masks"], out["e2e_masks_dilated"] = fm, md
        pcfg = RI.ProPainterConfig(e["ref_stride"], e["neighbor_length"], e["subvideo_length"], e["raft_iter"],
                                   "disable", e["T"], torch.device("cpu"), icfg.process_size)
        from reference.utils.model_utils import Models
        models = Models(raft, rfc, gen)
        uf, um, pf = RI.process_inpainting(models, ft, fm, md, pcfg)
        comp = RI.feature_propagation(gen, uf, um, md, pf, orig, pcfg)
        out["e2e_updated_frames"] = uf
        out["e2e_pred_flow_f"] = pf[0]
        out["e2e_frames_u8"] = torch.from_numpy(np.stack(comp))
    store = {}
    for k, v in out.items():
        a = v.detach().cpu().numpy()
        store[k] = a if a.dtype == np.uint8 else a.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **store)
    for k, v in store.items():
        print(k, v.shape, v.dtype, float(np.abs(v.astype(np.float64)).mean()))


if __name__ == "__main__":
    main()
