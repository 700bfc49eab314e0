"""CPU oracle for the HICom compressor (``mm_projector.forward``).

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (torch CPU tensors, no
CUDA, no custom ops) of the reference algorithm in
``/root/reference/hicom/model/projector.py`` and ``hicom/mm_utils.py:92-140``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it — as the checker or the reported CPU
baseline, never as the product path.  The product (``hicom_b200``) never imports
``oracle``; it fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against the reference module itself, executed in the
authoring container through ``oracle/ref_shim.py``:
  * ``tests/test_oracle_vs_reference.py`` runs both on seeded inputs (skipped when
    ``/root/reference`` is not mounted);
  * ``oracle/make_golden.py`` stored the reference's outputs in ``tests/golden/``;
    ``tests/test_oracle_golden.py`` checks the oracle against those everywhere.

Every function cites the reference lines it follows.  The op ORDER follows the
reference (so a bf16 run rounds where the reference rounds); the code is written
functionally over a flat ``state_dict`` instead of ``nn.Module`` classes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-6  # projector.py:318,403,565  partial(nn.LayerNorm, eps=1e-6)


# --------------------------------------------------------------------------------------
# type-string mini-language                                       projector.py:231-304
# --------------------------------------------------------------------------------------
@dataclass
class LocalSpec:
    temporal_kernel: int
    spatial_kernel: int
    adapt_q: bool = False
    adapt_k: bool = False
    adapt_v: bool = False
    adapt_guide: bool = False
    force_use_guide: object = False  # False or a mode string


@dataclass
class GlobalSpec:
    num_queries: int
    use_pos_emb: bool = True
    adapt_guide: bool = False
    force_use_guide: object = False


@dataclass
class ProjectorSpec:
    kind: str  # "mlp", "linear" or "hicom"
    mlp_depth: int = 0
    local: Optional[LocalSpec] = None
    global_: Optional[GlobalSpec] = None


def _leading_digits(text: str) -> str:
    out = ""
    for ch in text:
        if not ch.isdigit():
            break
        out += ch
    return out


def parse_projector_type(projector_type: str) -> ProjectorSpec:
    """projector.py:231-304.  Anything the grammar does not name (e.g. ``_coarse``) is ignored."""
    import re

    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)  # :233
    if m:
        return ProjectorSpec("mlp", mlp_depth=int(m.group(1)))
    if projector_type == "linear":  # :242
        return ProjectorSpec("linear")

    spec = ProjectorSpec("hicom")
    if "local" in projector_type:  # :247-282
        phase = projector_type.split("local")[-1].split("global")[0]
        digits = _leading_digits(phase)
        tk = int(digits[0])  # IndexError on empty, like the reference (:255)
        if len(digits) == 2:
            sk = int(digits[1])
        elif len(digits) == 3:
            sk = int(digits[1:3])
        else:
            # the reference leaves spatial_kernel_size unbound here -> UnboundLocalError (:256-259,280)
            raise UnboundLocalError("spatial_kernel_size")
        flags = dict(q=False, k=False, v=False, g=False)
        if "adapt" in phase:  # :262-273
            for ch in phase.split("adapt")[-1]:
                if ch in flags:
                    flags[ch] = True
                else:
                    break
        force = False
        if "guide" in phase:  # :276-277
            force = phase.split("guide")[-1].split("_")[0]
        spec.local = LocalSpec(tk, sk, flags["q"], flags["k"], flags["v"], flags["g"], force)
    if "global" in projector_type:  # :284-302
        phase = projector_type.split("global")[-1].split("local")[0]
        nq = int(_leading_digits(phase))
        force = False
        if "guide" in phase:
            force = phase.split("guide")[-1].split("_")[0]
        spec.global_ = GlobalSpec(nq, True, "adaptg" in phase, force)
    return spec


# --------------------------------------------------------------------------------------
# 3-D sincos position table                                        projector.py:57-101
# --------------------------------------------------------------------------------------
def sincos_axis_table(n: int, d_model: int) -> np.ndarray:
    """One axis of projector.py:70-93: full-width table, sin on even channels, cos on odd.

    The divisor uses ``np.float32(d_model)`` exactly as the reference (:71); the rest is float64.
    """
    pos = np.arange(n)[:, None]
    i = np.arange(d_model)[None, :]
    ang = pos / np.power(10000, (2 * (i // 2)) / np.float32(d_model))
    tab = np.zeros_like(ang)
    tab[:, 0::2] = np.sin(ang[:, 0::2])
    tab[:, 1::2] = np.cos(ang[:, 1::2])
    return tab


def pos_embed_3d(t: int, h: int, w: int, d_model: int) -> Tensor:
    """projector.py:95-101 + :606 — (t,h,w,d) float32 = f(t)+f(h)+f(w), summed in float64."""
    pt = sincos_axis_table(t, d_model)[:, None, None, :]
    ph = sincos_axis_table(h, d_model)[None, :, None, :]
    pw = sincos_axis_table(w, d_model)[None, None, :, :]
    return torch.from_numpy(pt + ph + pw).float()


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def _has(sd: Dict[str, Tensor], key: str) -> bool:
    return key in sd


def mlp(sd, prefix: str, x: Tensor) -> Tensor:
    """build_mlp, projector.py:307-312: Linear, then (GELU-erf, Linear) pairs at indices 2,4,…"""
    y = F.linear(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"])
    idx = 2
    while f"{prefix}.{idx}.weight" in sd:
        y = F.gelu(y)
        y = F.linear(y, sd[f"{prefix}.{idx}.weight"], sd[f"{prefix}.{idx}.bias"])
        idx += 2
    return y


def layer_norm(sd, prefix: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], LN_EPS)


def mha(sd, prefix: str, query: Tensor, key: Tensor, value: Tensor, num_heads: int,
        logit_scale: Optional[Tensor] = None, logit_bias: Optional[Tensor] = None) -> Tensor:
    """MultiheadAttention.forward, projector.py:166-228 (batch x len x channel inputs)."""
    b, ql, dim = query.shape
    kl = key.shape[1]
    hd = dim // num_heads
    q = F.linear(query, sd[f"{prefix}.q_proj.weight"], sd[f"{prefix}.q_proj.bias"])  # :180
    k = F.linear(key, sd[f"{prefix}.k_proj.weight"], sd[f"{prefix}.k_proj.bias"])  # :181
    v = F.linear(value, sd[f"{prefix}.v_proj.weight"], sd[f"{prefix}.v_proj.bias"])  # :182
    if logit_scale is not None:  # :184-188  L2 norm over the FULL width, before the head split
        q = q / q.norm(p=2, dim=-1, keepdim=True)
        k = k / k.norm(p=2, dim=-1, keepdim=True)
        scale, bias = logit_scale.exp(), logit_bias
    else:
        scale, bias = hd ** -0.5, 0.0  # :145,190-191
    q = q.view(b, ql, num_heads, hd).transpose(1, 2)
    k = k.view(b, kl, num_heads, hd).transpose(1, 2)
    v = v.view(b, kl, num_heads, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * scale + bias  # :197
    p = F.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype)  # :213
    o = torch.matmul(p, v)  # :215
    o = o.transpose(1, 2).contiguous().reshape(b, ql, dim)  # :223-224
    return F.linear(o, sd[f"{prefix}.out_proj.weight"], sd[f"{prefix}.out_proj.bias"])  # :226


def _adapt_guide(sd, prefix: str, g: Tensor) -> Tensor:
    """projector.py:364-365 / :388-389.  text2qk_proj is Identity for SigLIP (1152 == 1152)."""
    if _has(sd, f"{prefix}.text2qk_proj.0.weight"):
        g = mlp(sd, f"{prefix}.text2qk_proj", g)
    if _has(sd, f"{prefix}.guide_alpha"):
        a = sd[f"{prefix}.guide_alpha"]
        g = (1 - a) * g + a * layer_norm(sd, f"{prefix}.guide_norm", mlp(sd, f"{prefix}.guide_proj", g))
    return g


def guide_inject(mode, sd, prefix: str, visual: Tensor, guide: Optional[Tensor]) -> Tensor:
    """GuideInjector.forward, projector.py:344-397; IdentityMap (:104-110) for None/"off"."""
    if mode in (None, "off"):
        return visual
    if mode in ("direct", "coarse"):  # :352-372
        if visual.ndim == 4:
            t, h, w = visual.shape[:3]
            g = guide.reshape(1, 1, 1, -1).repeat(t, h, w, 1)  # guide must be rank 1 (:355)
        elif visual.ndim == 2:
            g = guide.reshape(1, -1).repeat(visual.shape[0], 1)
        else:
            raise ValueError("Invalid input shape for guide embedding.")
        if guide.ndim != 1:
            raise ValueError("direct/coarse guide must be a (d,) vector")  # einops 'd -> …' fails
        g = _adapt_guide(sd, prefix, g)
        if mode == "direct":
            return g  # :367-368 — the visual query content is discarded
        film = mlp(sd, f"{prefix}.coarse_proj", g)  # :370
        scale, shift = torch.chunk(film, 2, dim=-1)  # :371
        return layer_norm(sd, f"{prefix}.coarse_norm", visual * (1 + scale) + shift)  # :372
    if mode == "fine":  # :374-397
        if guide.ndim != 2:
            raise ValueError("fine guide must be (L, d) tokens")
        if visual.ndim == 4:
            t, h, w = visual.shape[:3]
            q = visual.reshape(t * h * w, 1, -1)
            g = guide.unsqueeze(0).repeat(t * h * w, 1, 1)
        elif visual.ndim == 2:
            q = visual.unsqueeze(0)
            g = guide.unsqueeze(0)
        else:
            raise ValueError("Invalid input shape for guide embedding.")
        g = _adapt_guide(sd, prefix, g)
        heads = visual.shape[-1] // 128  # :341
        a = mha(sd, f"{prefix}.fine_proj", q, g, g, heads)
        out = layer_norm(sd, f"{prefix}.fine_norm", q + a)  # :392
        return out.reshape(visual.shape)
    raise NotImplementedError(mode)  # :350


# --------------------------------------------------------------------------------------
# window gather                                                   projector.py:473-522
# --------------------------------------------------------------------------------------
def balanced_window_starts(n: int, k: int):
    """Start/stop of each window along one axis — projector.py:501-522.

    ``ceil(n/k)`` windows; the first ``n % split`` (or all, if that is 0) have length ``k`` taken
    fresh, the others take ``k-1`` fresh elements and step back by one so every window that can
    has length ``k``.  Reproduces the reference's quirks, including unequal lengths for some n
    (e.g. n in {5,6,9}, k=4), which make its ``torch.stack`` raise.
    """
    split = math.ceil(n / k)
    keep = n % split
    if keep == 0:
        keep = split
    spans = []
    start = 0
    for i in range(split):
        fresh = k - (0 if i < keep else 1)
        stop = start + fresh
        if fresh < k:
            start -= 1
        spans.append((start, stop))
        start = stop
    return spans


def _axis_windows(x: Tensor, k: int) -> Tensor:
    """(n, …) -> (k', n_windows, …): members first, like the reference's 't2 t1 …' layout."""
    n = x.shape[0]
    if n % k == 0:
        return x.reshape(n // k, k, *x.shape[1:]).transpose(0, 1)  # '(t1 t2) … -> t2 t1 …' (:477)
    pieces = [x[a:b] for a, b in balanced_window_starts(n, k)]
    return torch.stack(pieces, dim=1)  # raises RuntimeError on unequal pieces, like :520


def window_gather(x: Tensor, kernel) -> Tensor:
    """divide_feature, projector.py:473-499: (t,h,w,d) -> (t1·h1·w1, t2·h2·w2, d)."""
    kt, kh, kw = kernel
    y = _axis_windows(x, kt)  # t2 t1 h w d
    y = y.permute(2, 0, 1, 3, 4)  # h t2 t1 w d   (:481)
    y = _axis_windows(y, kh)  # h2 h1 t2 t1 w d
    y = y.permute(4, 0, 1, 2, 3, 5)  # w h2 h1 t2 t1 d   (:487)
    y = _axis_windows(y, kw)  # w2 w1 h2 h1 t2 t1 d
    w2, w1, h2, h1, t2, t1, d = y.shape
    y = y.permute(5, 3, 1, 4, 2, 0, 6)  # t1 h1 w1 t2 h2 w2 d   (:493)
    return y.reshape(t1 * h1 * w1, t2 * h2 * w2, d)


# --------------------------------------------------------------------------------------
# local compressor                                                projector.py:524-559
# --------------------------------------------------------------------------------------
def _mix(sd, prefix: str, name: str, x: Tensor) -> Tensor:
    """(1-α)x + α·LN(MLP(x)) — projector.py:533-534,541.  α absent ⇒ α == 0 but the reference
    still evaluates the expression with Identity modules, i.e. 1·x + 0·x (two passes)."""
    key = f"{prefix}.{name}_alpha"
    if key in sd:
        a = sd[key]
        if name == "q":  # q_proj is a bias-free Linear (:433)
            y = F.linear(x, sd[f"{prefix}.q_proj.weight"])
        else:
            y = mlp(sd, f"{prefix}.{name}_proj", x)
        return (1 - a) * x + a * layer_norm(sd, f"{prefix}.{name}_norm", y)
    return (1 - 0) * x + 0 * x


def local_compress(spec: LocalSpec, mode, sd, prefix: str, X: Tensor, E: Optional[Tensor],
                   guide: Optional[Tensor], modal: str, logit_scale=None, logit_bias=None,
                   qk_dim: int = 1152) -> Tensor:
    t, h, w = X.shape[:3]
    if E is not None and logit_scale is not None:  # :527-529
        E = E / E.norm(p=2, dim=-1, keepdim=True)
        guide = guide / guide.norm(p=2, dim=-1, keepdim=True)
    E = X if E is None else E  # :532
    key = _mix(sd, prefix, "k", E)  # :533
    value = _mix(sd, prefix, "v", X)  # :534
    tk = 1 if (modal == "image" or t == 1) else spec.temporal_kernel  # :536
    sk = spec.spatial_kernel
    ds = (math.ceil(t / tk), math.ceil(h / sk), math.ceil(w / sk))  # :537
    q = F.interpolate(X.permute(3, 0, 1, 2).unsqueeze(0), size=ds, mode="trilinear")  # :539
    q = q.squeeze(0).permute(1, 2, 3, 0)  # :540
    q = _mix(sd, prefix, "q", q)  # :541
    query = guide_inject(mode, sd, f"{prefix}.guide_injector", q, guide)  # :542
    rk = window_gather(key, (tk, sk, sk))  # :544
    rv = window_gather(value, (tk, sk, sk))  # :545
    rq = window_gather(query, (1, 1, 1))  # :546
    s = torch.bmm(rq, rk.permute(0, 2, 1))
    if logit_scale is not None:
        a = torch.softmax(s * logit_scale.exp() + logit_bias, dim=-1)  # :549
    else:
        a = torch.softmax(s / math.sqrt(qk_dim), dim=-1)  # :551
    o = torch.bmm(a, rv)  # :553
    o = o.reshape(ds[0], ds[1], ds[2], -1)  # :554-558 (t2=h2=w2=1)
    return mlp(sd, f"{prefix}.readout", o)  # :559


# --------------------------------------------------------------------------------------
# global compressor                                               projector.py:634-646
# --------------------------------------------------------------------------------------
def global_compress(spec: GlobalSpec, mode, sd, prefix: str, X: Tensor, guide: Optional[Tensor],
                    logit_scale=None, logit_bias=None, t0: int = 0) -> Tensor:
    """``t0`` is the oracle-side hook for frame shards (rows t0… of the position table, A8)."""
    t, h, w = X.shape[:3]
    if spec.use_pos_emb:  # :636-640
        pe = pos_embed_3d(t0 + t, h, w, X.shape[-1])[t0:].to(X.dtype)
        X = X + pe
    query = guide_inject(mode, sd, f"{prefix}.guide_injector", sd[f"{prefix}.query"], guide)  # :642
    heads = X.shape[-1] // 128  # :579
    kv = X.reshape(1, t * h * w, -1)
    a = mha(sd, f"{prefix}.attn_layer", query.unsqueeze(0), kv, kv, heads, logit_scale, logit_bias)  # :645
    return mlp(sd, f"{prefix}.readout", query + a.squeeze(0))  # :646


# --------------------------------------------------------------------------------------
# token layout                                                     mm_utils.py:92-140
# --------------------------------------------------------------------------------------
def post_process(merge_type: str, newline_position: str, feat: Tensor, modal: str,
                 image_newline: Optional[Tensor], is_anyres: bool) -> Tensor:
    d = feat.shape[-1]
    if merge_type.startswith("spatial") and merge_type != "flat":
        if modal == "video":
            t, h, w = feat.shape[:3]
            if newline_position == "grid":  # :101-107 newline after every row of every frame
                nl = image_newline.to(feat.device).expand(t, h, 1, d)
                return torch.cat([feat, nl], dim=2).reshape(-1, d)
            if newline_position == "frame":  # :108-114 newline after every frame
                nl = image_newline.to(feat.device).expand(t, 1, d)
                return torch.cat([feat.reshape(t, h * w, d), nl], dim=1).reshape(-1, d)
            if newline_position == "one_token":  # :115-117
                return torch.cat([feat.reshape(-1, d), image_newline[None].to(feat.device)], dim=0)
            if newline_position == "no_token":  # :118-119
                return feat.reshape(-1, d)
            raise ValueError(f"Unexpected mm_newline_position: {newline_position}")
        if modal == "image":
            if is_anyres:  # :124-130 newline after every row
                _, h, w, _ = feat.shape
                nl = image_newline.to(feat.device).expand(h, 1, d)
                return torch.cat([feat[0], nl], dim=1).reshape(-1, d)
            if image_newline is not None:  # :131-133
                return torch.cat([feat.reshape(-1, d), image_newline[None].to(feat.device)], dim=0)
            return feat.reshape(-1, d)  # :134-135
        return feat  # the reference falls through untouched for other modal strings
    return feat.reshape(-1, d)  # :96-97,137-138


# --------------------------------------------------------------------------------------
# the projector                                                   projector.py:649-708
# --------------------------------------------------------------------------------------
@dataclass
class OracleProjector:
    """Functional stand-in for ``HIComProjector``: a parsed spec + config knobs + a state_dict."""

    projector_type: str
    use_guide: object = None
    merge_type: str = "flat"
    newline_position: str = "one_token"
    state: Dict[str, Tensor] = field(default_factory=dict)
    local_logit: Optional[tuple] = None  # (logit_scale, logit_bias) tensors when use_clip_scale has 'local'
    global_logit: Optional[tuple] = None
    qk_dim: int = 1152  # 768 for the CLIP-L tower (projector.py:407-414)

    def __post_init__(self):
        self.spec = parse_projector_type(self.projector_type)
        assert self.spec.kind == "hicom"
        assert self.spec.local is not None or self.spec.global_ is not None  # :674

    def _mode(self, sub) -> object:
        mode = self.use_guide if sub.force_use_guide is False else sub.force_use_guide  # :422,585
        return mode

    def _local(self, X, E, g, modal):
        ls, lb = self.local_logit if self.local_logit else (None, None)
        return local_compress(self.spec.local, self._mode(self.spec.local), self.state,
                              "local_compressor", X, E, g, modal, ls, lb, qk_dim=self.qk_dim)

    def _global(self, X, g, t0=0):
        ls, lb = self.global_logit if self.global_logit else (None, None)
        return global_compress(self.spec.global_, self._mode(self.spec.global_), self.state,
                               "global_compressor", X, g, ls, lb, t0=t0)

    def forward(self, frames_feature, frames_embed, guide_embed, modal, image_newline=None) -> Tensor:
        pp = lambda x, anyres: post_process(self.merge_type, self.newline_position, x, modal,
                                            image_newline, anyres)
        local_x = global_x = None
        if self.spec.local is not None:  # :678-692
            if isinstance(frames_feature, dict):
                base = None
                if frames_feature["base"] is not None:
                    bx = frames_feature["base"].unsqueeze(0)
                    be = frames_embed["base"].unsqueeze(0) if frames_embed is not None else None
                    base = pp(self._local(bx, be, guide_embed, modal), False)
                px = frames_feature["patch"].unsqueeze(0)
                pe = frames_embed["patch"].unsqueeze(0) if frames_embed is not None else None
                patch = pp(self._local(px, pe, guide_embed, modal), True)
                local_x = torch.cat([base, patch], dim=-2) if base is not None else patch
            else:
                local_x = pp(self._local(frames_feature, frames_embed, guide_embed, modal), False)
        if self.spec.global_ is not None:  # :694-700
            if isinstance(frames_feature, dict):
                global_x = self._global(frames_feature["patch"].unsqueeze(0), guide_embed)
            else:
                global_x = self._global(frames_feature, guide_embed)
        if local_x is None:
            return global_x
        if global_x is None:
            return local_x
        return torch.cat([local_x, global_x], dim=-2)  # :707

    __call__ = forward


# --------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs shared by fixtures, tests and bench
# --------------------------------------------------------------------------------------
def param_shapes(projector_type: str, use_guide, hidden: int, d: int = 1152) -> Dict[str, tuple]:
    """Names and shapes of the reference state_dict for a configuration (A3, A6, A7)."""
    spec = parse_projector_type(projector_type)
    shapes: Dict[str, tuple] = {}

    def add_mlp(prefix, i, o):
        shapes[f"{prefix}.0.weight"] = (o, i)
        shapes[f"{prefix}.0.bias"] = (o,)
        shapes[f"{prefix}.2.weight"] = (o, o)
        shapes[f"{prefix}.2.bias"] = (o,)

    def add_ln(prefix, n):
        shapes[f"{prefix}.weight"] = (n,)
        shapes[f"{prefix}.bias"] = (n,)

    def add_injector(prefix, mode, adapt_guide):
        if mode in (None, "off"):
            return
        if adapt_guide:
            add_mlp(f"{prefix}.guide_proj", d, d)
            add_ln(f"{prefix}.guide_norm", d)
            shapes[f"{prefix}.guide_alpha"] = (1,)
        if mode == "coarse":
            add_mlp(f"{prefix}.coarse_proj", d, 2 * d)
            add_ln(f"{prefix}.coarse_norm", d)
        elif mode == "fine":
            for p in ("k_proj", "v_proj", "q_proj", "out_proj"):
                shapes[f"{prefix}.fine_proj.{p}.weight"] = (d, d)
                shapes[f"{prefix}.fine_proj.{p}.bias"] = (d,)
            add_ln(f"{prefix}.fine_norm", d)

    if spec.local is not None:
        lp = "local_compressor"
        mode = use_guide if spec.local.force_use_guide is False else spec.local.force_use_guide
        add_injector(f"{lp}.guide_injector", mode, spec.local.adapt_guide)
        if spec.local.adapt_q and mode != "direct":
            shapes[f"{lp}.q_proj.weight"] = (d, d)
            add_ln(f"{lp}.q_norm", d)
            shapes[f"{lp}.q_alpha"] = (1,)
        if spec.local.adapt_k:
            add_mlp(f"{lp}.k_proj", d, d)
            add_ln(f"{lp}.k_norm", d)
            shapes[f"{lp}.k_alpha"] = (1,)
        if spec.local.adapt_v:
            add_mlp(f"{lp}.v_proj", d, d)
            add_ln(f"{lp}.v_norm", d)
            shapes[f"{lp}.v_alpha"] = (1,)
        add_mlp(f"{lp}.readout", d, hidden)
    if spec.global_ is not None:
        gp = "global_compressor"
        mode = use_guide if spec.global_.force_use_guide is False else spec.global_.force_use_guide
        shapes[f"{gp}.query"] = (spec.global_.num_queries, d)
        add_injector(f"{gp}.guide_injector", mode, spec.global_.adapt_guide)
        for p in ("k_proj", "v_proj", "q_proj", "out_proj"):
            shapes[f"{gp}.attn_layer.{p}.weight"] = (d, d)
            shapes[f"{gp}.attn_layer.{p}.bias"] = (d,)
        add_mlp(f"{gp}.readout", d, hidden)
    return shapes


def synth_state_dict(projector_type: str, use_guide, hidden: int, seed: int = 0,
                     dtype=torch.float32, d: int = 1152) -> Dict[str, Tensor]:
    """Seeded weights independent of module construction order (fixtures regenerate them).

    Linear weights N(0, .02²) — the reference uses trunc_normal_(std=.02) (:157,464,625); biases
    N(0, .02²) instead of the reference's zeros so bias handling is exercised; LayerNorm γ = 1 +
    N(0,.1²), β = N(0,.1²); ``query`` N(0,.02²) (zero-init in the reference, :583); ``*_alpha`` = 0.5.
    Values are drawn in fp32 in sorted-name order and cast once.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in sorted(param_shapes(projector_type, use_guide, hidden, d).items()):
        if name.endswith("_alpha"):
            v = torch.full(shape, 0.5)
        elif "norm.weight" in name:
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "norm.bias" in name:
            v = 0.1 * torch.randn(shape, generator=g)
        else:
            v = 0.02 * torch.randn(shape, generator=g)
        sd[name] = v.to(dtype)
    return sd


def synth_inputs(T: int, H: int, W: int, guide_kind: Optional[str], seed: int, L: int = 32,
                 d: int = 1152, std: float = 0.5, dtype=torch.float32):
    """Seeded (frames_feature, frames_embed, guide_embed) — SURVEY §8d: N(0, 0.5²), fp32 then cast once.

    ``guide_kind``: None (no guide: frames_embed and guide are None), "vec" (d,), "tokens" (L,d).
    """
    g = torch.Generator().manual_seed(seed)
    X = (std * torch.randn(T, H, W, d, generator=g)).to(dtype)
    if guide_kind is None:
        return X, None, None
    E = (std * torch.randn(T, H, W, d, generator=g)).to(dtype)
    if guide_kind == "vec":
        gd = (std * torch.randn(d, generator=g)).to(dtype)
    else:
        gd = (std * torch.randn(L, d, generator=g)).to(dtype)
    return X, E, gd


def guide_kind_for(use_guide) -> Optional[str]:
    if use_guide in (None, "off"):
        return None
    return "tokens" if use_guide == "fine" else "vec"


def rel_err(a: Tensor, b: Tensor) -> float:
    """max|a-b| / max|b| — the tolerance definition of SURVEY §8c."""
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def cosine(a: Tensor, b: Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
