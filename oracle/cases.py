"""Parity cases shared by the golden generator, the CPU tests and the GPU tests (test infrastructure).

Each case regenerates its weights and inputs from seeds (``hicom_oracle.synth_state_dict`` /
``synth_inputs``), so fixtures only store the reference OUTPUT.  Shapes follow SURVEY.md §8c's test
hygiene list: all guide modes, adapt variants, local-only / global-only, T in {1,4,7,8,16}, image
modal, newline layouts, non-divisible (balanced-window) grids, plus BASELINE config c1 at full size.
"""
from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class Case:
    name: str
    ptype: str
    use_guide: Optional[str]
    T: int
    H: int = 6
    W: int = 6
    hidden: int = 64
    dtype: str = "float32"
    modal: str = "video"
    merge: str = "flat"
    nlpos: str = "one_token"
    newline: bool = False
    wseed: int = 1
    xseed: int = 5


CASES = [
    # BASELINE.json configs[0]: width 896, 16 frames x 729 x 1152, fp32 (guide mode per SURVEY §8d headline)
    Case("c1_coarse_896_T16", "local43_global32", "coarse", 16, 27, 27, 896),
    Case("c1_direct_896_T16", "local43_global32_coarse", "direct", 16, 27, 27, 896),
    # guide modes, small grids
    Case("none_T8", "local43_global32", None, 8),
    Case("direct_T8", "local43_global32", "direct", 8),
    Case("coarse_T8", "local43_global32", "coarse", 8),
    Case("fine_T8", "local43_global32", "fine", 8),
    Case("off_T4", "local43_global32", "off", 4),
    # frame counts incl. T==1 (temporal kernel 1) and the balanced overlapping windows (T=7)
    Case("coarse_T1", "local43_global32", "coarse", 1),
    Case("coarse_T4", "local43_global32", "coarse", 4),
    Case("coarse_T7", "local43_global32", "coarse", 7),
    Case("none_T2_short_window", "local43_global32", None, 2),
    Case("coarse_nondiv_7x8", "local43_global32", "coarse", 7, 7, 8),
    Case("coarse_27x27_T4", "local43_global32", "coarse", 4, 27, 27, 128),
    # adapters and structural variants
    Case("adaptkv_coarse_T8", "local43_adaptkv_global32", "coarse", 8),
    Case("adaptqkvg_fine_T8", "local43_adaptqkvg_global32_adaptg", "fine", 8),
    Case("adaptqkvg_coarse_T4", "local43_adaptqkvg_global32_adaptg", "coarse", 4),
    Case("global_only_coarse_T8", "global32", "coarse", 8),
    Case("local_only_fine_T8", "local43", "fine", 8),
    Case("local22_global8_T8", "local22_global8", "coarse", 8),
    Case("forced_guide", "local43guidecoarse_global32guidedirect", None, 4),
    # image modal + newline layouts (mm_utils.py:99-135)
    Case("image_T1_newline", "local43_global32", "coarse", 1, 6, 6, 64, "float32", "image", "spatial_unpad", "one_token", True),
    Case("video_grid_newline", "local43_global32", "coarse", 8, 6, 6, 64, "float32", "video", "spatial_unpad", "grid", True),
    Case("video_frame_newline", "local43_global32", "coarse", 8, 6, 6, 64, "float32", "video", "spatial_unpad", "frame", True),
    Case("video_one_token", "local43_global32", "direct", 8, 6, 6, 64, "float32", "video", "spatial_unpad", "one_token", True),
    Case("video_no_token", "local43_global32", "direct", 8, 6, 6, 64, "float32", "video", "spatial_unpad", "no_token", True),
    # the reference's own bf16 forward (stored for information; gates use the fp32 truth)
    Case("bf16_coarse_T8", "local43_global32", "coarse", 8, 9, 9, 128, "bfloat16"),
    Case("bf16_none_T8", "local43_global32", None, 8, 9, 9, 128, "bfloat16"),
    Case("bf16_fine_T8", "local43_global32", "fine", 8, 9, 9, 128, "bfloat16"),
]

CASES_BY_NAME = {c.name: c for c in CASES}


def materialise(case: Case):
    """Return (state_dict fp32, X, E, guide, newline) for a case; tensors on CPU, in the case dtype."""
    import torch

    from . import hicom_oracle as O

    dt = getattr(torch, case.dtype)
    sd = O.synth_state_dict(case.ptype, case.use_guide, case.hidden, seed=case.wseed, dtype=dt)
    spec = O.parse_projector_type(case.ptype)
    kinds = []
    for sub in (spec.local, spec.global_):
        if sub is not None:
            mode = case.use_guide if sub.force_use_guide is False else sub.force_use_guide
            kinds.append(O.guide_kind_for(mode))
    kind = next((k for k in kinds if k is not None), None)
    X, E, g = O.synth_inputs(case.T, case.H, case.W, kind, seed=case.xseed, dtype=dt)
    nl = None
    if case.newline:
        gen = torch.Generator().manual_seed(case.xseed + 77)
        nl = (0.02 * torch.randn(case.hidden, generator=gen)).to(dt)
    return sd, X, E, g, nl
