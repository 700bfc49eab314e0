"""Generate tests/golden/reference_outputs.pt from the REAL reference (authoring container only).

    python -m oracle.make_golden

For each case in ``oracle/cases.py`` the unmodified reference module
(``/root/reference/hicom/model/projector.py`` via ``oracle/ref_shim.py``) is built from a plain config,
loaded with the seeded synthetic state_dict (strict=True) and run on the seeded inputs.  Only the
outputs are stored (fp32 copies; bf16 values are exactly representable), plus torch/numpy versions.
"""
import os
import sys

import numpy
import torch

from . import hicom_oracle as O
from .cases import CASES, materialise
from .ref_shim import BagConfig, load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_outputs.pt")


def run_reference(ref, case):
    sd, X, E, g, nl = materialise(case)
    cfg = BagConfig(use_guide=case.use_guide, hidden_size=case.hidden, max_num_frames=4,
                    mm_projector_type=case.ptype, mm_patch_merge_type=case.merge,
                    mm_newline_position=case.nlpos)
    m = ref.build_vision_projector(cfg)
    m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    m = m.to(getattr(torch, case.dtype)).eval()
    with torch.no_grad():
        return m(X, E, g, case.modal, nl)


def main():
    ref = load_reference()
    if ref is None:
        sys.exit("/root/reference is not mounted; golden vectors can only be made in the authoring container")
    torch.set_num_threads(os.cpu_count())
    blob = {"_meta": {"torch": str(torch.__version__), "numpy": str(numpy.__version__),
                      "source": "hicom/model/projector.py:676-708 via oracle/ref_shim.py"}}
    for case in CASES:
        out = run_reference(ref, case)
        blob[case.name] = out.float().clone()
        print(f"{case.name:28s} {tuple(out.shape)} {case.dtype}")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(blob, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
