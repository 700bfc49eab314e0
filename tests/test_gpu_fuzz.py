"""GPU: seeded random shapes through forward_batched against the oracle (odd tile counts, ragged grids, tiny and
non-divisible frame counts, every guide mode, both dtypes).  Same tolerances as test_gpu_parity.py."""
import random

import pytest
import torch

from oracle import hicom_oracle as O

from util import Cfg, to_dev

pytestmark = pytest.mark.gpu


def _cases(n=28, seed=7, dtypes=("float32", "bfloat16")):
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        T = rng.choice([1, 2, 3, 4, 7, 8, 10, 12, 16, 20])          # 5, 6, 9 raise in the reference (tested elsewhere)
        H, W = rng.randint(3, 13), rng.randint(3, 13)
        B = rng.choice([1, 2, 3, 5])
        if T * H * W * B > 7000:
            continue
        out.append((T, H, W, B, rng.choice([None, "direct", "coarse", "fine"]), rng.choice([64, 128, 896]),
                    rng.choice(list(dtypes)), rng.choice(["local43_global32", "local43_global32",
                                                                     "local22_global8", "local43_global5"]), len(out)))
    return out


# fp16 (the reference's inference dtype) gets its own seeded set so that the fp32 / bf16 case list stays what it was
@pytest.mark.parametrize("T,H,W,B,guide,hidden,dtype,ptype,idx", _cases() + _cases(14, seed=11, dtypes=("float16",)),
                         ids=lambda v: str(v) if not isinstance(v, str) else v)
def test_random_shapes_match_oracle(T, H, W, B, guide, hidden, dtype, ptype, idx, built_library):
    dt = getattr(torch, dtype)
    from hicom_b200.projector import build_vision_projector
    sd = O.synth_state_dict(ptype, guide, hidden, seed=100 + idx, dtype=dt)
    vids = [O.synth_inputs(T, H, W, O.guide_kind_for(guide), seed=1000 + 10 * idx + b, dtype=dt) for b in range(B)]
    orc = O.OracleProjector(ptype, guide, state={k: v.float() for k, v in sd.items()})
    f = lambda t: None if t is None else t.float()
    try:
        with torch.no_grad():
            want = torch.stack([orc.forward(f(x), f(e), f(g), "video") for x, e, g in vids])
    except RuntimeError:
        want = None  # the reference cannot stack unequal windows for this grid: the CUDA path must refuse it too
    m = build_vision_projector(Cfg(use_guide=guide, hidden_size=hidden, mm_projector_type=ptype, max_num_frames=4))
    m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    m = m.to(dt).cuda().eval()
    X = torch.stack([v[0] for v in vids])
    E = None if vids[0][1] is None else torch.stack([v[1] for v in vids])
    G = None if vids[0][2] is None else torch.stack([v[2] for v in vids])
    if want is None:
        with torch.no_grad(), pytest.raises(RuntimeError):
            m.forward_batched(to_dev(X), to_dev(E), to_dev(G), "video")
        return
    with torch.no_grad():
        got = m.forward_batched(to_dev(X), to_dev(E), to_dev(G), "video").float().cpu()
    assert got.shape == want.shape
    if dtype == "float32":
        assert O.rel_err(got, want) <= 1e-4
    elif dtype == "float16":
        assert O.cosine(got, want) >= 0.99999 and O.rel_err(got, want) <= 3e-3
    else:
        assert O.cosine(got, want) >= 0.999 and O.rel_err(got, want) <= 1e-2
