"""Flow-completion stage alone at the bench size (80 frames 640x360): timing for A/B switches (PP_PROG=0/1) and a
small target for ncu captures of the propagation-step kernels.

    python tools/rfc_bench.py [T H W reps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from comfyui_propainter_nodes_b200 import weights as Wt
from comfyui_propainter_nodes_b200.engine import Engine
from comfyui_propainter_nodes_b200.synthetic import synthetic_mask


def main():
    T, H, W, reps = [int(x) for x in (sys.argv[1:5] + [80, 360, 640, 5][len(sys.argv) - 1:])]
    dev = torch.device("cuda:0")
    eng = Engine(dev, workspace_gb=24.0).load_weights(Wt.synthetic_raft_state_dict(), Wt.synthetic_rfc_state_dict(),
                                                      Wt.synthetic_generator_state_dict())
    g = torch.Generator().manual_seed(0)
    ff = (torch.randn(T - 1, 2, H // 8, W // 8, generator=g) * 2).to(dev)
    ff = torch.nn.functional.interpolate(ff, size=(H, W), mode="bilinear") + 1.5
    fb = -ff
    masks = synthetic_mask(T, H, W)[:, None].contiguous().to(dev)
    for _ in range(2):
        eng.flow_complete(ff, fb, masks)
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.flow_complete(ff, fb, masks)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    l0 = eng.launch_count
    eng.flow_complete(ff, fb, masks)
    print(json.dumps({"T": T, "H": H, "W": W, "PP_PROG": os.environ.get("PP_PROG", "1"), "ms": sorted(times)[len(times) // 2],
                      "ms_all": [round(t, 2) for t in times], "launches": eng.launch_count - l0}))


if __name__ == "__main__":
    main()
