"""Shared test helpers: build the CUDA projector / the oracle for a parity case."""
import torch

from oracle import hicom_oracle as O
from oracle.cases import Case, materialise


class Cfg:
    """Attribute-bag config like the HF config object the reference reads."""

    def __init__(self, **kw):
        self.mm_vision_tower = "google/siglip-so400m-patch14-384"
        self.mm_hidden_size = 1152
        self.hidden_size = 896
        self.mm_projector_type = "local43_global32"
        for k, v in kw.items():
            setattr(self, k, v)


def cfg_for(case: Case) -> Cfg:
    return Cfg(use_guide=case.use_guide, hidden_size=case.hidden, max_num_frames=4,
               mm_projector_type=case.ptype, mm_patch_merge_type=case.merge, mm_newline_position=case.nlpos)


def oracle_for(case: Case, sd, dtype=None) -> O.OracleProjector:
    if dtype is not None:
        sd = {k: v.to(dtype) for k, v in sd.items()}
    return O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, sd)


def cuda_module_for(case: Case, sd, device="cuda"):
    import hicom_b200
    m = hicom_b200.build_vision_projector(cfg_for(case))
    m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    return m.to(getattr(torch, case.dtype)).to(device).eval()


def to_dev(x, device="cuda"):
    return None if x is None else x.to(device)


def truth_fp32(case: Case):
    """fp32 oracle on the case's (possibly bf16-rounded) weights and inputs — the bf16 'truth' (SURVEY §8c)."""
    sd, X, E, g, nl = materialise(case)
    f = lambda t: None if t is None else t.float()
    orc = oracle_for(case, sd, torch.float32)
    with torch.no_grad():
        return orc.forward(f(X), f(E), f(g), case.modal, f(nl))
