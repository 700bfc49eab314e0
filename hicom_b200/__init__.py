"""hicom_b200 — B200-native HICom video-token compressor (drop-in for hicom/model/projector.py)."""
from .projector import (GlobalCompressor, GuideInjector, HIComProjector, IdentityMap, LocalCompressor,  # noqa: F401
                        MultiheadAttention, build_mlp, build_vision_projector, get_3d_position_embedding,
                        load_mm_projector)

__version__ = "0.1.0"
