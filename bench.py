#!/usr/bin/env python
"""Bench of the ProPainter hot path: inpainted frames/s at 640x360 on an 80-frame subvideo (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the whole hot path (RAFT -> flow completion -> image propagation -> sliding-window
generator -> composite) over one synthetic 80-frame 640x360 clip per GPU.  `value` is measured with the
prepared tensors already resident in HBM; `e2e` goes through the ComfyUI node call with host tensors (host
pre-processing, H2D, D2H of the result inside the timed region).  Weights are seeded synthetic checkpoints of
the real architectures (no network in this environment).
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

T_FRAMES, HEIGHT, WIDTH = 80, 360, 640
_OUT_FD = 1


def emit(line: str) -> None:
    """The ONE JSON line of the contract goes to the real stdout; everything else this process or its libraries
    print on fd 1 (NCCL's version banner, the node's progress lines) is routed to stderr in main()."""
    os.write(_OUT_FD, (line + "\n").encode())


PARAMS = dict(mask_dilates=5, flow_mask_dilates=8, ref_stride=10, neighbor_length=10, subvideo_length=80, raft_iter=20,
              fp16="enable")
METRIC = "inpainted frames/sec at 640x360, 80-frame subvideo"
WORKLOAD = "configs[1]: 80-frame 640x360 synthetic clip, ref_stride=10 neighbor_length=10 raft_iter=20 fp16, 1 clip per B200"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [int(s[0]) for s in self.samples if s and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=int(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def synthetic_inputs():
    from comfyui_propainter_nodes_b200.synthetic import synthetic_clip, synthetic_mask
    return synthetic_clip(T_FRAMES, HEIGHT, WIDTH, 1234), synthetic_mask(T_FRAMES, HEIGHT, WIDTH)


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference itself is a Python package that cannot travel to the GPU box)
# ------------------------------------------------------------------------------------------------------------------
def cpu_threads():
    """Host threads for the CPU arm: all cores up to 16 (beyond that the small 1/8-res convs of the recurrent
    stages get slower with more threads on the 128-core host; measured in profiles/)."""
    return min(os.cpu_count() or 1, int(os.environ.get("PP_CPU_THREADS", 16)))


def cpu_sample(n_frames=3):
    """Times the CPU oracle on a bounded sample of the same workload: the first `n_frames` frames of the clip with
    the workload's parameters.  Returns (frames/s, seconds)."""
    from comfyui_propainter_nodes_b200 import weights as Wt
    from comfyui_propainter_nodes_b200.utils import image_utils as IU
    from oracle import propainter_oracle as O
    image, mask = synthetic_inputs()
    image, mask = image[:n_frames], mask[:n_frames]
    icfg = IU.ImageConfig(WIDTH, HEIGHT, PARAMS["mask_dilates"], PARAMS["flow_mask_dilates"], (WIDTH, HEIGHT), n_frames)
    ft, fm, md, orig = IU.prepare_frames_and_masks(IU.convert_image_to_frames(image), mask, icfg, torch.device("cpu"))
    sds = (Wt.synthetic_raft_state_dict(), Wt.synthetic_rfc_state_dict(), Wt.synthetic_generator_state_dict())
    t0 = time.perf_counter()
    O.run_pipeline(*sds, ft, fm, md, orig, raft_iter=PARAMS["raft_iter"], subvideo_length=PARAMS["subvideo_length"],
                   neighbor_length=PARAMS["neighbor_length"], ref_stride=PARAMS["ref_stride"])
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    cores = cpu_threads()
    torch.set_num_threads(cores)
    n = 3
    for _ in range(min(args.warmup, 1)):
        cpu_sample(n)
    vals = []
    for _ in range(args.steps):
        v, _ = cpu_sample(n)
        vals.append(v)
    v = float(np.mean(vals))
    sample = f"first {n} frames of the 80-frame 640x360 clip, raft_iter=20, fp32, CPU oracle port of the reference"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * n / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world):
    import torch.distributed as dist
    from comfyui_propainter_nodes_b200 import weights as Wt
    from comfyui_propainter_nodes_b200 import propainter_inference as PI
    from comfyui_propainter_nodes_b200.propainter_nodes import ProPainterInpaint, _to_host
    from comfyui_propainter_nodes_b200.utils import image_utils as IU
    from comfyui_propainter_nodes_b200.utils import model_utils as MU

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    models = MU.build_models(dev, Wt.synthetic_raft_state_dict(), Wt.synthetic_rfc_state_dict(),
                             Wt.synthetic_generator_state_dict(), workspace_gb=64.0)
    MU._CACHE[str(dev)] = models          # the node's initialize_models() finds the resident engine
    eng = models.raft_model.engine
    image, mask = synthetic_inputs()      # every rank processes its own (identical) 80-frame subvideo: weak scaling
    icfg = IU.ImageConfig(WIDTH, HEIGHT, PARAMS["mask_dilates"], PARAMS["flow_mask_dilates"], (WIDTH, HEIGHT), T_FRAMES)
    ft, fm, md, orig = IU.prepare_frames_and_masks(IU.convert_image_to_frames(image), mask, icfg, dev)
    orig_dev = torch.from_numpy(np.stack(orig)).to(dev)
    cfg = PI.ProPainterConfig(PARAMS["ref_stride"], PARAMS["neighbor_length"], PARAMS["subvideo_length"],
                              PARAMS["raft_iter"], PARAMS["fp16"], T_FRAMES, dev, icfg.process_size)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB)

    strong = args.mode == "strong" and world > 1

    def step():
        if strong:
            from comfyui_propainter_nodes_b200.parallel import inpaint_clip_distributed
            return inpaint_clip_distributed(models, ft, fm, md, orig_dev, cfg)
        uf, um, flows = PI.process_inpainting(models, ft, fm, md, cfg)
        return PI.feature_propagation_device(models.inpaint_model, uf, um, md, flows, orig_dev, cfg)

    def staged():
        """Same work as step(), with CUDA events between the stages (reported as stage_ms)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        gt = PI.compute_flow(models.raft_model, ft, cfg)
        ev[1].record()
        pf = PI.complete_flow(models.flow_model, gt, fm, cfg.subvideo_length)
        ev[2].record()
        uf, um = PI.image_propagation(models.inpaint_model, ft, md, pf, cfg)
        ev[3].record()
        PI.feature_propagation_device(models.inpaint_model, uf, um, md, pf, orig_dev, cfg)
        ev[4].record()
        torch.cuda.synchronize()
        names = ["raft", "flow_completion", "image_propagation", "generator_windows"]
        return {n: round(ev[i].elapsed_time(ev[i + 1]), 2) for i, n in enumerate(names)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = eng.launch_count
    times = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = step()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    launches = (eng.launch_count - l0) // max(args.steps, 1)
    barrier()
    total_ms = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = (1 if strong else world) * T_FRAMES / (ms_per_step / 1000.0)

    # ---- end to end through the node API with host tensors (pre-processing + H2D + D2H inside)
    node = ProPainterInpaint()
    img_host, mask_host = image.pin_memory(), mask.pin_memory()
    def e2e_step():
        if strong:   # host tensors -> device pre-processing -> sharded clip -> float IMAGE back on the host
            from comfyui_propainter_nodes_b200.parallel import inpaint_clip_distributed
            f, m1, m2, o = eng.preprocess(img_host, mask_host, PARAMS["flow_mask_dilates"], PARAMS["mask_dilates"])
            return _to_host(eng.postprocess(inpaint_clip_distributed(models, f, m1, m2, o, cfg)))
        with contextlib.redirect_stdout(sys.stderr):   # the node prints progress; stdout carries only the JSON line
            frames, _, _ = node.propainter_inpainting(img_host, mask_host, WIDTH, HEIGHT, **PARAMS)
        return frames
    e2e_value = None
    if not args.no_e2e:
        # two warm-up calls whose results are held the way a caller (ComfyUI's output cache) holds them: the IMAGE
        # result lives in page-locked blocks of torch's caching host allocator, and the steady state alternates
        # between two blocks (the previous result is released only after the next one exists)
        res = e2e_step()
        res = e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            res = e2e_step()
        torch.cuda.synchronize()
        e2e_s = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = (1 if strong else world) * T_FRAMES / float(e2e_s.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = img_host.numel() * 4 + mask_host.numel() * 4      # the node uploads the IMAGE / MASK float tensors
    d2h = T_FRAMES * HEIGHT * WIDTH * 3 * 4                  # and downloads the float32 IMAGE result

    # ---- per-kernel timing of one extra step (CUDA events on the launch stream) for the roofline
    roof, extra, stage_ms = None, [], None
    if strong and not args.no_profile and rank != 0:
        step()                      # the profiled step below is collective in strong mode
    if rank == 0 and not args.no_profile:
        stage_ms = staged()
        pk = peaks()
        eng.profile_enable(True)
        step()
        prof = eng.profile_dump()
        eng.profile_enable(False)
        conv = {k: v for k, v in prof.items() if k.startswith("conv:")}
        conv_ms = sum(v["ms"] for v in conv.values())
        conv_fl = sum(v["flops"] for v in conv.values())
        all_ms = sum(v["ms"] for v in prof.values())
        ach = conv_fl / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "conv_halo_kernel + conv_igemm_kernel (tcgen05 convolutions: TMA halo-tile kernel for stride-1 convs and linears, cp.async implicit GEMM for the rest), aggregate over all conv launches of the step",
                "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                "traffic": None, "peak_source": pk["source"] + " bf16 sustained", "share_of_profiled_step": conv_ms / all_ms if all_ms else None,
                "launches": sum(v["count"] for v in conv.values())}
        for name in ("corr_lookup", "imgprop_step", "dcn_sample", "featprop_warp", "fold_ffn"):
            if name in prof and prof[name]["ms"] > 0:
                v = prof[name]
                gbs = v["bytes"] / (v["ms"] / 1e3) / 1e9
                extra.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                              "frac": gbs / pk["hbm"], "launches": v["count"], "ms": v["ms"]})
        if "attention" in prof:
            v = prof["attention"]
            extra.append({"kernel": "window_attention_tc (tcgen05) + window_attention (mma.sync, unmasked windows); flops are an upper bound (all windows masked)", "bound": "tensor",
                          "ms": v["ms"], "launches": v["count"]})
        if args.profile_out:
            with open(args.profile_out, "w") as fh:
                json.dump({"stage_ms": stage_ms, "kernels": prof}, fh, indent=1, sort_keys=True)
        top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]
        extra.append({"top_by_time_ms": {k: round(v["ms"], 3) for k, v in top}, "profiled_step_kernel_ms": all_ms})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = cpu_threads()
        torch.set_num_threads(cores)
        v, dt = cpu_sample(3)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"first 3 frames of the same clip and parameters through the CPU oracle ({dt:.1f} s)"}
    if rank == 0:
        emit(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": T_FRAMES / world if strong else T_FRAMES, "l2": "flushed between steps (256 MiB write)",
                       "weights": "seeded synthetic checkpoints", "parallelism": (f"1 subvideo sharded over {world} GPUs (RAFT pairs + windows, 2 all-gathers)" if strong
                                       else f"{world} independent subvideos, no data-path collective"),
                       "roofline_timing": "one extra profiled step after the timed region"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "roofline_other": extra,
            "stage_ms": stage_ms, "cpu_baseline": cpu, "workspace_peak_gb": eng.workspace_peak / 2 ** 30,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = one 80-frame subvideo per GPU (default); strong = ONE subvideo shared by all GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true", help="skip the node-level end-to-end leg (profiling runs)")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel timed extra step")
    ap.add_argument("--profile-out", default=None, help="write the full per-kernel table (JSON) here")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    global _OUT_FD
    sys.stdout.flush()
    _OUT_FD = os.dup(1)      # keep the real stdout for the JSON line ...
    os.dup2(2, 1)            # ... and send every other write to fd 1 (C libraries included) to stderr
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world)


if __name__ == "__main__":
    main()
