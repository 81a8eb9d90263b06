"""Model residency: three checkpoints -> one packed sm_100a engine (reference: utils/model_utils.py:13-59).

``Models`` keeps the reference's three fields; each is a thin stage handle sharing one ``Engine``.
Checkpoints are the reference's ``.pth`` state_dict files under ``<package>/weights/``
(raft-things.pth, recurrent_flow_completion.pth, ProPainter.pth).  Unlike the reference, the engine is
cached per device instead of being rebuilt on every node execution.  There is no download step here
(no network); place the files, or pass state dicts to ``build_models``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights")
CHECKPOINTS = ("raft-things.pth", "recurrent_flow_completion.pth", "ProPainter.pth")


class StageHandle:
    """What the orchestration functions receive where the reference passes an nn.Module."""

    def __init__(self, engine, stage: str):
        self.engine = engine
        self.stage = stage

    def __repr__(self):
        return f"<propainter_b200 stage '{self.stage}' on {self.engine.device}>"


@dataclass
class Models:
    raft_model: StageHandle
    flow_model: StageHandle
    inpaint_model: StageHandle


_CACHE = {}


def build_models(device, raft_sd, rfc_sd, gen_sd, workspace_gb: float = 48.0) -> Models:
    from ..engine import Engine
    eng = Engine(device, workspace_gb=workspace_gb).load_weights(raft_sd, rfc_sd, gen_sd)
    return Models(StageHandle(eng, "raft"), StageHandle(eng, "flow_completion"), StageHandle(eng, "inpaint"))


def initialize_models(device: torch.device, use_half: str = "enable") -> Models:
    """Load the three checkpoints from ``weights/`` (same files as the reference) into a cached engine."""
    key = str(device)
    if key in _CACHE:
        return _CACHE[key]
    paths = [os.path.join(WEIGHTS_DIR, n) for n in CHECKPOINTS]
    missing = [p for p in paths if not os.path.exists(p)]
    if missing:
        raise FileNotFoundError(
            "ProPainter checkpoints not found: " + ", ".join(missing) +
            " (download raft-things.pth, recurrent_flow_completion.pth and ProPainter.pth from the "
            "sczhou/ProPainter v0.1.0 release into the weights/ directory)")
    sds = [torch.load(p, map_location="cpu") for p in paths]
    _CACHE[key] = build_models(device, *sds)
    return _CACHE[key]
