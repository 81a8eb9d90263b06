"""Where the node-level (host tensors in, host tensor out) time goes: H2D + device pre-processing, the hot path,
post-processing + D2H.  Synchronises between phases, so the sum is slightly above the pipelined e2e number."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from comfyui_propainter_nodes_b200 import weights as Wt, propainter_inference as PI
from comfyui_propainter_nodes_b200.utils import image_utils as IU, model_utils as MU
from comfyui_propainter_nodes_b200.propainter_nodes import ProPainterInpaint, _to_host


def main():
    dev = torch.device("cuda", 0)
    models = MU.build_models(dev, Wt.synthetic_raft_state_dict(), Wt.synthetic_rfc_state_dict(),
                             Wt.synthetic_generator_state_dict(), workspace_gb=64.0)
    eng = models.raft_model.engine
    MU.set_resident_models(dev, models)   # the node's initialize_models() returns the resident engine
    image, mask = B.synthetic_inputs()
    image, mask = image.pin_memory(), mask.pin_memory()
    P = B.PARAMS
    icfg = IU.ImageConfig(B.WIDTH, B.HEIGHT, P["mask_dilates"], P["flow_mask_dilates"], (B.WIDTH, B.HEIGHT), B.T_FRAMES)
    cfg = PI.ProPainterConfig(P["ref_stride"], P["neighbor_length"], P["subvideo_length"], P["raft_iter"], P["fp16"],
                              B.T_FRAMES, dev, icfg.process_size)

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for it in range(4):
        t0 = sync()
        ft, fm, md, orig = eng.preprocess(image, mask, P["flow_mask_dilates"], P["mask_dilates"])
        t1 = sync()
        uf, um, flows = PI.process_inpainting(models, ft, fm, md, cfg)
        t2 = sync()
        comp = PI.feature_propagation_device(models.inpaint_model, uf, um, md, flows, orig, cfg)
        t3 = sync()
        out = eng.postprocess(comp)
        t4 = sync()
        host = _to_host(out)
        t5 = sync()
        u8 = comp.cpu()
        t6 = sync()
        host2 = torch.div(u8, 255.0)
        t7 = time.perf_counter()
        print(f"iter {it}: preprocess+H2D {1e3*(t1-t0):.1f}  process_inpainting {1e3*(t2-t1):.1f}  feature_propagation "
              f"{1e3*(t3-t2):.1f}  postprocess {1e3*(t4-t3):.1f}  D2H float32 pinned(_to_host) {1e3*(t5-t4):.1f} | alt: D2H uint8 "
              f"{1e3*(t6-t5):.1f} + host /255 {1e3*(t7-t6):.1f}  equal={bool(torch.equal(host, host2))} threads={torch.get_num_threads()}",
              flush=True)


def node_loop():
    import contextlib
    dev = torch.device("cuda", 0)
    node = ProPainterInpaint()
    image, mask = B.synthetic_inputs()
    image, mask = image.pin_memory(), mask.pin_memory()
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):
            res = node.propainter_inpainting(image, mask, B.WIDTH, B.HEIGHT, **B.PARAMS)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"iter node {it}: {1e3*(t1-t0):.1f} ms", flush=True)
        del res


if __name__ == "__main__":
    main()
    node_loop()
