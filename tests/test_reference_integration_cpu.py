"""Host-side drop-in checks that need no GPU.

* The drop-in classes go through the REFERENCE's own module-swap loops (validate_mlm_config.py:174-191,
  validate_clm.py:147-168: construct with the script's kwargs, ``load_state_dict(old.state_dict(), strict=False)``) and
  through the reference's exact-type ``specials`` dispatch (quantization/autoquant_utils.py:236-270, table at
  quantized_bert.py:633-637) -- run against the live tree when /root/reference is present (build container), skipped on
  the GPU box.
* Regression tests for the round-1 review: symmetric weight-quantizer checkpoints reload into a fresh model,
  ``gate_scaling_factor`` multiplies conditional gates only, hooked S / P probes force the unfused path, an ``nn.GELU``
  FFN activation is fused into the QuantLinear (tap after the activation).
"""
import os
import sys

import pytest
import torch
from torch import nn

REF = "/root/reference/OutEffHop"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")


@pytest.fixture(scope="module")
def ref_tree():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    G._install_shims()
    yield
    for p in (REF,):
        if p in sys.path:
            sys.path.remove(p)


def _tiny_bert():
    from transformers import BertConfig, BertForMaskedLM
    cfg = BertConfig(vocab_size=97, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                     max_position_embeddings=64)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    return cfg, BertForMaskedLM(cfg).eval()


@needs_ref
@pytest.mark.parametrize("gate", ["none", "conditional_per_token", "unconditional_per_head"])
def test_bert_swap_loop_of_validate_script_with_drop_in_classes(ref_tree, gate):
    """validate_mlm_config.py:174-191 verbatim, with the three imports pointed at this package; the result must carry the
    same parameters (names, shapes, values) as the same loop run with the reference's own class."""
    from transformers_language.models import bert_attention as RB, softmax as RS
    from outeffhop_b200.bert_attention import AttentionGateType, BertSelfAttentionWithExtras
    from outeffhop_b200.softmax import SOFTMAX_MAPPING

    assert list(SOFTMAX_MAPPING) == list(RS.SOFTMAX_MAPPING)                       # argparse choices= (args.py:345-351)
    assert AttentionGateType.list_names() == list(RB.AttentionGateType.__members__)
    args = dict(attn_softmax="clippedsoftmax1(-.025:1)", alpha=None, max_seq_length=32, skip_attn=False,
                attn_gate_type=gate, attn_gate_init=0.25, attn_gate_mlp=False, attn_gate_mlp2=False,
                attn_gate_linear_all_features=False)

    def swap(model, cls, mapping, gate_enum):
        for layer_idx in range(len(model.bert.encoder.layer)):
            old_self = model.bert.encoder.layer[layer_idx].attention.self
            torch.manual_seed(100 + layer_idx)                                     # same gate init on both sides
            new_self = cls(model.config, softmax_fn=mapping[args["attn_softmax"]], alpha=args["alpha"],
                           max_seq_length=args["max_seq_length"], skip_attn=args["skip_attn"],
                           attn_gate_type=gate_enum[args["attn_gate_type"]], attn_gate_init=args["attn_gate_init"],
                           attn_gate_mlp=args["attn_gate_mlp"], attn_gate_mlp2=args["attn_gate_mlp2"],
                           attn_gate_linear_all_features=args["attn_gate_linear_all_features"])
            new_self.load_state_dict(old_self.state_dict(), strict=False)
            model.bert.encoder.layer[layer_idx].attention.self = new_self
        return model

    _, ours = _tiny_bert()
    _, theirs = _tiny_bert()
    swap(ours, BertSelfAttentionWithExtras, SOFTMAX_MAPPING, AttentionGateType)
    swap(theirs, RB.BertSelfAttentionWithExtras, RS.SOFTMAX_MAPPING, RB.AttentionGateType)
    sd_o, sd_t = ours.state_dict(), theirs.state_dict()
    assert list(sd_o) == list(sd_t)
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_t[k]), k
    # a checkpoint written by the reference model loads strictly into the drop-in model and vice versa
    ours.load_state_dict(sd_t, strict=True)
    theirs.load_state_dict(sd_o, strict=True)
    # attributes the quantized wrappers copy (quantized_bert.py:227-258)
    a, b = ours.bert.encoder.layer[0].attention.self, theirs.bert.encoder.layer[0].attention.self
    for name in ("num_attention_heads", "attention_head_size", "all_head_size", "position_embedding_type", "is_decoder",
                 "attn_gate_init", "attn_gate_mlp", "attn_gate_mlp2", "attn_gate_linear_all_features"):
        assert getattr(a, name) == getattr(b, name), name
    assert a.attn_gate_type.name == b.attn_gate_type.name
    assert [n for n, _ in a.named_modules()] == [n for n, _ in b.named_modules()]


@needs_ref
def test_opt_swap_loop_of_validate_script_with_drop_in_classes(ref_tree):
    """validate_clm.py:147-168 with the drop-in class."""
    from transformers import OPTConfig, OPTForCausalLM
    from transformers_language.models import opt_attention as RO, softmax as RS, bert_attention as RB
    from outeffhop_b200.bert_attention import AttentionGateType
    from outeffhop_b200.opt_attention import OPTAttentionWithExtras
    from outeffhop_b200.softmax import SOFTMAX_MAPPING

    def build(cls, mapping, gate_enum):
        cfg = OPTConfig(vocab_size=101, hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2,
                        word_embed_proj_dim=128, max_position_embeddings=64, do_layer_norm_before=True)
        cfg._attn_implementation = "eager"
        torch.manual_seed(1)
        model = OPTForCausalLM(cfg).eval()
        for layer_idx in range(len(model.model.decoder.layers)):
            old_attn = model.model.decoder.layers[layer_idx].self_attn
            torch.manual_seed(200 + layer_idx)
            new_attn = cls(embed_dim=old_attn.embed_dim, num_heads=old_attn.num_heads, dropout=old_attn.dropout,
                           is_decoder=True, bias=True, softmax_fn=mapping["softmax1"], alpha=None,
                           max_seq_length=32, skip_attn=False, attn_gate_type=gate_enum["conditional_per_token"],
                           attn_gate_init=0.25, attn_gate_mlp=False, attn_gate_mlp2=False,
                           attn_gate_linear_all_features=False)
            new_attn.load_state_dict(old_attn.state_dict(), strict=False)
            model.model.decoder.layers[layer_idx].self_attn = new_attn
        return model

    ours = build(OPTAttentionWithExtras, SOFTMAX_MAPPING, AttentionGateType)
    theirs = build(RO.OPTAttentionWithExtras, RS.SOFTMAX_MAPPING, RB.AttentionGateType)
    sd_o, sd_t = ours.state_dict(), theirs.state_dict()
    assert list(sd_o) == list(sd_t)
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_t[k]), k
    a, b = ours.model.decoder.layers[0].self_attn, theirs.model.decoder.layers[0].self_attn
    for name in ("embed_dim", "num_heads", "dropout", "head_dim", "scaling", "is_decoder"):
        assert getattr(a, name) == getattr(b, name), name


@needs_ref
def test_reference_quantize_model_dispatches_on_the_registered_drop_in_type(ref_tree):
    """autoquant_utils.py:244 picks the wrapper by ``type(model) in specials`` -- a subclass of the reference class would
    not match, so the drop-in class is registered with its own wrapper.  The reference's
    recursive ``quantize_model`` must find it through an HF BertAttention container and leave the Linear / LayerNorm
    children to its own module_map."""
    from quantization.autoquant_utils import quantize_model
    from transformers_language.quant_configs import get_quant_config
    from transformers_language.utils import val_qparams
    from outeffhop_b200 import hf_compat
    from outeffhop_b200.bert_attention import BertSelfAttentionWithExtras
    from outeffhop_b200.quantized_attention import QuantizedBertSelfAttentionWithExtras
    from outeffhop_b200 import quantized_models as QM

    from outeffhop_b200.softmax import SOFTMAX_MAPPING
    cfg, model = _tiny_bert()
    old_self = model.bert.encoder.layer[0].attention.self
    core = BertSelfAttentionWithExtras(cfg, softmax_fn=SOFTMAX_MAPPING["softmax1"], max_seq_length=32)   # the script's loop
    core.load_state_dict(old_self.state_dict(), strict=False)
    # (the HF >= 4.48 call-convention adapter of hf_compat is a SUBCLASS: it needs its own entry in the table)
    assert type(hf_compat.swap_bert_attention(model, attn_softmax="softmax1", max_seq_length=32)
                .bert.encoder.layer[0].attention.self) is not BertSelfAttentionWithExtras
    seen = {}

    def wrapper(org, **quant_params):                      # the registration a maintainer adds to quantized_bert.py:633-637
        seen["kw"] = sorted(quant_params)
        return QuantizedBertSelfAttentionWithExtras(org, **QM.val_qparams(QM.get_quant_config()))

    holder = nn.ModuleDict({"self": core})
    qparams = val_qparams(get_quant_config())
    q = quantize_model(holder, specials={BertSelfAttentionWithExtras: wrapper}, **qparams)
    assert type(q["self"]) is QuantizedBertSelfAttentionWithExtras
    assert "act_method" in seen["kw"] and "n_bits" in seen["kw"]
    # the wrapper exposes the reference wrapper's tap names (quantized_bert.py:242-244) and quantized projections
    names = {n for n, _ in q["self"].named_modules()}
    for n in ("attn_scores_act_quantizer", "attn_probs_act_quantizer", "context_act_quantizer", "query", "key", "value"):
        assert n in names, n


# ---------------------------------------------------------------------------------------------------------------------
def test_symmetric_quantizer_checkpoint_reloads_into_fresh_instance():
    from outeffhop_b200.quantization import QuantizationManager, SymmetricUniformQuantizer, CurrentMinMaxEstimator
    torch.manual_seed(0)
    w = torch.randn(16, 8)
    m = QuantizationManager(qmethod=SymmetricUniformQuantizer, init=CurrentMinMaxEstimator, qparams=dict(n_bits=8))
    m.range_estimator.forward = lambda x: (x.min(), x.max())               # CPU: the estimator kernels are CUDA-only
    wq = m(w)
    sd = m.state_dict()
    assert "quantizer._delta" in sd and "quantizer._signed" in sd
    fresh = QuantizationManager(qmethod=SymmetricUniformQuantizer, init=CurrentMinMaxEstimator, qparams=dict(n_bits=8))
    fresh.load_state_dict(sd, strict=True)                                 # raised "Unexpected key(s)" before the fix
    assert fresh.quantizer.is_initialized and fresh.quantizer.signed
    fresh.fix_ranges()
    assert torch.equal(fresh(w), wq)


def test_gate_scaling_factor_only_multiplies_conditional_gates():
    """bert_attention.py:295-297 vs :327; opt_attention.py:277-279 vs :309."""
    from transformers import BertConfig
    from outeffhop_b200.bert_attention import AttentionGateType, BertSelfAttentionWithExtras
    from outeffhop_b200.opt_attention import OPTAttentionWithExtras
    cfg = BertConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256)
    for gt, want in ((AttentionGateType.unconditional_per_head, 1.0), (AttentionGateType.conditional_per_token, 4.0),
                     (AttentionGateType.conditional_per_head, 4.0)):
        m = BertSelfAttentionWithExtras(cfg, attn_gate_type=gt, attn_gate_init=0.25, fine_tuning=True)
        assert m.gate_scaling_factor == 4.0 and m._gate_scale() == want, gt
        o = OPTAttentionWithExtras(embed_dim=128, num_heads=2, attn_gate_type=gt, attn_gate_init=0.25, fine_tuning=True)
        assert o.gate_scaling_factor == 4.0 and o.effective_gate_scale == want, gt


def test_hooked_probes_are_detected():
    from transformers import BertConfig
    from outeffhop_b200.bert_attention import BertSelfAttentionWithExtras
    m = BertSelfAttentionWithExtras(BertConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256))
    assert not m._probes_hooked()
    h = m.attn_probs_before_dropout.register_forward_hook(lambda *a: None)
    assert m._probes_hooked()
    h.remove()
    assert not m._probes_hooked()


def test_nn_gelu_is_fused_into_the_ffn_quant_linear():
    """quantize_intermediate -> quantize_sequential: activations_set members run inside the QuantLinear, before its tap
    (autoquant_utils.py:105-176); HF's GELUActivation stays a separate module behind the tapped dense output."""
    from outeffhop_b200 import hf_compat, quantized_models as QM
    from outeffhop_b200.quantization import QuantLinear
    _, model = _tiny_bert()
    hf_compat.swap_bert_attention(model, attn_softmax="softmax1", max_seq_length=32)
    q = QM.quantize_bert_for_masked_lm(model)
    inter = q.bert.encoder.layer[0].intermediate
    assert len(inter) == 2 and isinstance(inter[0], QuantLinear) and inter[0].activation_function is None
    _, model = _tiny_bert()
    for layer in model.bert.encoder.layer:
        layer.intermediate.intermediate_act_fn = nn.GELU()                   # validate_mlm_config.py:166-168
    hf_compat.swap_bert_attention(model, attn_softmax="softmax1", max_seq_length=32)
    q = QM.quantize_bert_for_masked_lm(model)
    inter = q.bert.encoder.layer[0].intermediate
    assert len(inter) == 1 and isinstance(inter[0].activation_function, nn.GELU)
    _, model = _tiny_bert()
    hf_compat.fuse_bert_gelu(hf_compat.swap_bert_attention(model, attn_softmax="softmax1", max_seq_length=32))
    q = QM.quantize_bert_for_masked_lm(model)
    inter = q.bert.encoder.layer[0].intermediate
    assert len(inter) == 1 and isinstance(inter[0].activation_function, hf_compat.GELUFused)
