"""B200-native ProPainter inference path behind the ComfyUI node API of daniabib/ComfyUI_ProPainter_Nodes."""
from .propainter_nodes import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
