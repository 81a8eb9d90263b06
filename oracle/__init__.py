"""CPU oracle for the ProPainter hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product path (comfyui_propainter_nodes_b200) never does.
"""
