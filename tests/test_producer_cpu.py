"""CPU: the producer-side oracle (oracle/siglip_head.py) pinned against HF transformers' own SigLIP classes driven by
the reference's call sequence (encoder.py:284-286), and the host logic of hicom_b200.producer."""
import pytest
import torch

from oracle import siglip_head as SH


def _hf_head(hidden=1152, inter=4304, heads=16):
    tf = pytest.importorskip("transformers")
    from transformers.models.siglip.modeling_siglip import SiglipMultiheadAttentionPoolingHead
    cfg = tf.SiglipVisionConfig(hidden_size=hidden, intermediate_size=inter, num_attention_heads=heads,
                                num_hidden_layers=1, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
    torch.manual_seed(0)
    head = SiglipMultiheadAttentionPoolingHead(cfg).eval()
    with torch.no_grad():
        for k, v in SH.synth_head_state(0, hidden, inter).items():
            head.state_dict()[k].copy_(v)
    return head


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 0.0), (torch.bfloat16, 0.0)])
def test_oracle_matches_transformers_head(dtype, tol):
    head = _hf_head().to(dtype)
    h = SH.synth_hidden(2, 36, seed=1).to(dtype)
    with torch.no_grad():
        # the reference's own lines, encoder.py:284-286
        ref = head.layernorm(h)
        ref = h + head.mlp(ref)
        ref = ref.reshape(2, 6, 6, -1)
        sd = {k: v for k, v in head.state_dict().items() if k.startswith(("layernorm.", "mlp."))}
        got = SH.image_embeds(sd, h, side=6)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert float((got.float() - ref.float()).abs().max()) <= tol


def test_default_config_is_so400m():
    """hidden_act / sizes the oracle assumes are what SiglipVisionConfig carries for so400m (README.md:20 pins 4.46.3;
    the defaults below are set explicitly by the checkpoint's config.json: 1152 / 4304 / gelu_pytorch_tanh / 1e-6)."""
    tf = pytest.importorskip("transformers")
    cfg = tf.SiglipVisionConfig()
    assert cfg.hidden_act == "gelu_pytorch_tanh" and abs(cfg.layer_norm_eps - 1e-6) < 1e-12


def test_module_parameter_names_and_sharing():
    from hicom_b200.producer import SiglipHeadEmbed
    head = _hf_head(hidden=128, inter=256, heads=2)
    m = SiglipHeadEmbed(128, 256)
    res = m.load_state_dict(head.state_dict(), strict=False)
    assert res.missing_keys == []                       # every parameter of ours exists in the HF head
    assert all(k.startswith(("probe", "attention.")) for k in res.unexpected_keys)
    shared = SiglipHeadEmbed.from_head(head)
    assert shared.mlp.fc1.weight is head.mlp.fc1.weight and shared.layernorm.bias is head.layernorm.bias
    assert shared.act == 2                               # HICOM_ACT_GELU_TANH


def test_no_cpu_fallback_and_bad_inputs():
    from hicom_b200.producer import SiglipHeadEmbed
    m = SiglipHeadEmbed(128, 256).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="CUDA"):
            m(torch.zeros(1, 4, 128))
        with pytest.raises(ValueError):
            m(torch.zeros(1, 5, 128))                    # 5 tokens are not a square grid
        with pytest.raises(TypeError):
            m(torch.zeros(1, 4, 128, dtype=torch.float64))       # fp32 / bf16 / fp16 are the storage types
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 4, 128))                        # autograd on: the training path, same refusal of CPU tensors
    from hicom_b200 import autograd as ag
    ag.enable(False)
    try:
        with pytest.raises(RuntimeError, match="forward-only"):
            m(torch.zeros(1, 4, 128))                    # training path switched off: fail loudly, never drop gradients
    finally:
        ag.enable(True)
    with pytest.raises(NotImplementedError):
        SiglipHeadEmbed(128, 256, layer_norm_eps=1e-5)
    with pytest.raises(NotImplementedError):
        SiglipHeadEmbed(128, 256, hidden_act="relu")
