"""Pins oracle/resvit_oracle.py against vectors produced by the UNMODIFIED reference res-vit/model.py
(tests/golden/resvit_tiny.pt) and, where /root/reference exists, against the live reference; checks the
product's Res-ViT constructors reproduce the reference state_dict (keys, shapes, init, frozen set).  CPU only."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

from conftest import grad_close, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, resvit_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "resvit_tiny.pt")


@pytest.mark.parametrize("variant", ["bs2", "bs1"])
def test_oracle_matches_reference_golden_train_and_eval(variant):
    g = torch.load(GOLD)[variant]
    args = SimpleNamespace(**g["args"])
    sd = {k: v.clone().requires_grad_(k in g["train"]["trainable"]) for k, v in g["state_dict"].items()}
    torch.manual_seed(g["gumbel_seed"])
    out = resvit_oracle.resvit_forward(sd, args, g["img"], g["labels"], training=True)
    t = g["train"]
    assert rel_l2(out["logits"], t["logits"]) < 1e-5
    assert torch.equal(out["acts"], t["acts"])            # hard routing decisions: bit-exact
    for name, key in (("c_loss", "c"), ("a_loss", "a"), ("d_loss", "d"), ("r_entropy", "e")):
        assert abs(float(out[name]) - float(t[key])) < 1e-5 * max(1.0, abs(float(t[key]))), name
    assert abs(out["active_metric"] - t["metric"]) < 1e-6
    (1.0 * out["c_loss"] + 2.0 * out["a_loss"] + 0.5 * out["d_loss"] + 0.1 * out["r_entropy"]).backward()
    for k, ref in t["grads"].items():
        assert sd[k].grad is not None, k
        assert grad_close(sd[k].grad, ref, 5e-5, atol=1e-8), (k, rel_l2(sd[k].grad, ref))
    with torch.no_grad():
        ev = resvit_oracle.resvit_forward({k: v.detach() for k, v in sd.items()}, args, g["img"], g["labels"], training=False)
    assert rel_l2(ev["logits"], g["eval"]["logits"]) < 1e-5
    assert torch.equal(ev["acts"], g["eval"]["acts"])
    assert abs(float(ev["r_entropy"]) - float(g["eval"]["e"])) < 1e-5


def test_lra_tables_match_survey_values():
    import vitb200
    from vitb200 import lra_tables
    for bs in (1, 2, 4):
        assert lra_tables.get_indices_from_LRA_mask(bs) == [tuple(map(list, r)) for r in resvit_oracle.lra_table(bs)]
    assert lra_tables.get_indices_from_LRA_mask(2) == [([1], [2, 3], [0]), ([0, 2], [1, 3], [])]
    with pytest.raises(ValueError):
        lra_tables.get_indices_from_LRA_mask(3)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")
def test_product_constructor_matches_live_reference_state_dict():
    import vitb200
    from vitb200 import resvit
    mod, mu = ref_loader.load_resvit_model()
    for bs in (1, 2, 4):
        assert [tuple(map(list, r)) for r in mu.get_indices_from_LRA_mask(bs)] == \
               [tuple(map(list, r)) for r in resvit_oracle.lra_table(bs)]
    kw = dict(dim=128, mlp_dim=256, n_layers=4, n_heads=2, n_kv_heads=2, dynamic_start_layer=1, dynamic_router_hdim=64,
              low_rank_dim=32, block_size=2, use_lora=True, use_reslr=True, image_size=(32, 32), patch_size=(16, 16),
              num_classes=7, device="cpu")
    torch.manual_seed(3)
    r = mod.Transformer(mod.ModelArgs(**kw))
    torch.manual_seed(3)
    m = resvit.Transformer(resvit.ModelArgs(**kw))
    rs, ms = r.state_dict(), m.state_dict()
    assert list(rs) == list(ms)
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    assert [n for n, p in r.named_parameters() if p.requires_grad] == [n for n, p in m.named_parameters() if p.requires_grad]
    m.load_state_dict(rs)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")
def test_plain_resvit_equals_src_vit_in_the_oracles():
    """with mapped weights the plain Res-ViT and the src ViT are the same function."""
    from oracle import vit_init, vit_oracle
    cfg = dict(image_size=(32, 32), patch_size=(16, 16), emb_dim=128, mlp_dim=256, num_heads=2, num_layers=2, num_classes=5)
    sd = vit_init.reference_state_dict(cfg, seed=4, scaled=True)
    m = {"cls_token": sd["cls_token"], "embedding.weight": sd["embedding.weight"], "embedding.bias": sd["embedding.bias"],
         "pos_embedding.pos_embedding": sd["transformer.pos_embedding.pos_embedding"],
         "norm.layer_norm.weight": sd["transformer.norm.weight"], "norm.layer_norm.bias": sd["transformer.norm.bias"],
         "classifier.weight": sd["classifier.weight"], "classifier.bias": sd["classifier.bias"]}
    for i in range(2):
        s, d = "transformer.encoder_layers.%d." % i, "layers.%d." % i
        for a, b in (("query", "wq"), ("key", "wk"), ("value", "wv")):
            m[d + "attention.%s.weight" % b] = sd[s + "attn.%s.weight" % a].reshape(128, 128).t().contiguous()
            m[d + "attention.%s.bias" % b] = sd[s + "attn.%s.bias" % a].reshape(128)
        m[d + "attention.wo.weight"] = sd[s + "attn.out.weight"].reshape(128, 128).t().contiguous()
        m[d + "attention.wo.bias"] = sd[s + "attn.out.bias"]
        for n1, n2 in (("norm1", "attention_norm"), ("norm2", "ffn_norm")):
            m[d + n2 + ".layer_norm.weight"], m[d + n2 + ".layer_norm.bias"] = sd[s + n1 + ".weight"], sd[s + n1 + ".bias"]
        for f in ("fc1", "fc2"):
            m[d + "feed_forward.%s.weight" % f], m[d + "feed_forward.%s.bias" % f] = sd[s + "mlp.%s.weight" % f], sd[s + "mlp.%s.bias" % f]
    args = SimpleNamespace(n_layers=2, n_heads=2, norm_eps=1e-5, use_lora=False, use_reslr=False, block_size=1,
                           dynamic_start_layer=2, dynamic_reserve_initials=1, dynamic_active_target=0.4)
    img = torch.randn(3, 3, 32, 32)
    labels = torch.tensor([0, 1, 2])
    out = resvit_oracle.resvit_forward(m, args, img, labels, training=False)
    assert rel_l2(out["logits"], vit_oracle.vit_logits(img, sd)) < 1e-6
