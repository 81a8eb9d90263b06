"""Run the other BASELINE.json configs once on one GPU through the node API (sanity + timing, no oracle):
C3 240-frame 640x360 (3 sub-video chunks), C4 80-frame 1280x720, C5 outpaint 640x360 -> 768x360."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contextlib
import torch

from comfyui_propainter_nodes_b200 import weights as Wt
from comfyui_propainter_nodes_b200.propainter_nodes import ProPainterInpaint, ProPainterOutpaint
from comfyui_propainter_nodes_b200.synthetic import synthetic_clip, synthetic_mask
from comfyui_propainter_nodes_b200.utils import model_utils as MU


def main():
    only = sys.argv[1:]
    dev = torch.device("cuda:0")
    models = MU.build_models(dev, Wt.synthetic_raft_state_dict(), Wt.synthetic_rfc_state_dict(),
                             Wt.synthetic_generator_state_dict(), workspace_gb=100.0)
    MU.set_resident_models(dev, models)   # the node's initialize_models() returns the resident engine
    common = dict(mask_dilates=5, flow_mask_dilates=8, ref_stride=10, neighbor_length=10, subvideo_length=80,
                  raft_iter=20, fp16="enable")
    cases = {
        "C3_240f_640x360": ("in", 240, 360, 640, {}),
        "C4_80f_1280x720": ("in", 80, 720, 1280, {}),
        "C5_outpaint_80f_640x360_x1.2": ("out", 80, 360, 640, dict(width_scale=1.2, height_scale=1.0)),
    }
    for name, (kind, T, H, W, extra) in cases.items():
        if only and not any(o in name for o in only):
            continue
        img = synthetic_clip(T, H, W, 77)
        res = None
        times = []
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr):
                if kind == "in":
                    res = ProPainterInpaint().propainter_inpainting(img, synthetic_mask(T, H, W), W, H, **common)
                else:
                    res = ProPainterOutpaint().propainter_outpainting(img, W, H, extra["width_scale"],
                                                                      extra["height_scale"], **common)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        out = res[0]
        eng = models.raft_model.engine
        rec = dict(config=name, frames=T, out_shape=list(out.shape), finite=bool(torch.isfinite(out).all()),
                   out_min=float(out.min()), out_max=float(out.max()), seconds=times, fps_e2e=T / times[-1],
                   workspace_peak_gb=eng.workspace_peak / 2 ** 30)
        if kind == "in":   # known pixels are returned untouched (quantised to uint8)
            md = res[2].cpu() > 0.5
            q = (img * 255).clamp(0, 255).to(torch.uint8).float() / 255.0
            rec["known_pixels_exact"] = bool(torch.equal(out[~md], q[~md]))
        else:
            rec["out_size"] = [int(res[2]), int(res[3])]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
