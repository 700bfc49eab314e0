"""Small end-to-end forward (fp32 + bf16, coarse and fine) for compute-sanitizer runs."""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle.cases import CASES_BY_NAME, materialise
from util import cuda_module_for

for name, dtype, T in (("coarse_27x27_T4", "bfloat16", 8), ("fine_T8", "bfloat16", 8), ("coarse_T7", "float32", 7)):
    case = dataclasses.replace(CASES_BY_NAME[name], dtype=dtype, T=T)
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    with torch.no_grad():
        out = m(X.cuda(), E.cuda(), g.cuda(), "video")
        outb = m.forward_batched(torch.stack([X, X]).cuda(), torch.stack([E, E]).cuda(), torch.stack([g, g]).cuda(), "video")
    torch.cuda.synchronize()
    print(name, dtype, tuple(out.shape), tuple(outb.shape), bool(torch.isfinite(out.float()).all()))
