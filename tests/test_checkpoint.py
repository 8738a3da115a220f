"""Checkpoint / weight-format row: JAX npz conversion, the reference's .pth layout, and the
src <-> Res-ViT state-dict mapping.  CPU only.  Where /root/reference exists the converters are compared with the
reference's own functions (its module-level `tensorflow` / `swanlab` imports are stubbed: they are only used to
open files / log); everywhere, the mapping is pinned by the derived identity SURVEY §8(c) names — a plain Res-ViT
computes exactly the src ViT's logits under the mapped weights."""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, resvit_oracle, vit_init, vit_oracle  # noqa: E402

CFG = dict(image_size=(32, 32), patch_size=(16, 16), emb_dim=64, mlp_dim=128, num_heads=4, num_layers=2, num_classes=10)


def _ckpt():
    import vitb200  # noqa: F401  (package import; the module itself is host-only)
    from vitb200 import checkpoint
    return checkpoint


def _fake_flax_tree(sd, heads):
    """The flax parameter tree a src state dict would have come from (inverse of the conversion rules)."""
    out = {}
    L = max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.encoder_layers.")) + 1
    out["cls"] = sd["cls_token"].numpy()
    out["embedding/kernel"] = sd["embedding.weight"].permute(2, 3, 1, 0).contiguous().numpy()
    out["embedding/bias"] = sd["embedding.bias"].numpy()
    out["Transformer/posembed_input/pos_embedding"] = sd["transformer.pos_embedding.pos_embedding"].numpy()
    out["Transformer/encoder_norm/scale"] = sd["transformer.norm.weight"].numpy()
    out["Transformer/encoder_norm/bias"] = sd["transformer.norm.bias"].numpy()
    out["head/kernel"] = sd["classifier.weight"].t().contiguous().numpy()
    out["head/bias"] = sd["classifier.bias"].numpy()
    for i in range(L):
        s, j = "transformer.encoder_layers.%d." % i, "Transformer/encoderblock_%d/" % i
        for n, ln in (("norm1", "LayerNorm_0"), ("norm2", "LayerNorm_2")):
            out[j + ln + "/scale"] = sd[s + n + ".weight"].numpy()
            out[j + ln + "/bias"] = sd[s + n + ".bias"].numpy()
        for n in ("query", "key", "value", "out"):
            out[j + "MultiHeadDotProductAttention_1/%s/kernel" % n] = sd[s + "attn.%s.weight" % n].numpy()
            out[j + "MultiHeadDotProductAttention_1/%s/bias" % n] = sd[s + "attn.%s.bias" % n].numpy()
        for n, dn in (("fc1", "Dense_0"), ("fc2", "Dense_1")):
            out[j + "MlpBlock_3/%s/kernel" % dn] = sd[s + "mlp.%s.weight" % n].t().contiguous().numpy()
            out[j + "MlpBlock_3/%s/bias" % dn] = sd[s + "mlp.%s.bias" % n].numpy()
    return out


def test_jax_npz_loads_into_the_src_state_dict(tmp_path):
    ck = _ckpt()
    sd = vit_init.reference_state_dict(CFG, seed=3, scaled=True)
    path = str(tmp_path / "ViT-tiny.npz")
    np.savez(path, **_fake_flax_tree(sd, CFG["num_heads"]))
    got = ck.load_checkpoint(path)
    assert set(got) == set(sd)
    for k in sd:
        assert got[k].shape == sd[k].shape and torch.equal(got[k], sd[k]), k
    # and the reference's own converted-checkpoint layout
    out = ck.save_jax_to_pytorch(path, str(tmp_path))
    again = ck.load_checkpoint(out)
    assert all(torch.equal(again[k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        ck.load_checkpoint(str(tmp_path / "weights.h5"))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")
def test_jax_conversion_equals_the_reference_converter():
    ck = _ckpt()
    tf = types.ModuleType("tensorflow")
    tfio = types.ModuleType("tensorflow.io")
    tfio.gfile = SimpleNamespace(GFile=open)
    tf.io = tfio
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "tensorflow.io")}
    sys.modules["tensorflow"], sys.modules["tensorflow.io"] = tf, tfio
    try:
        spec = importlib.util.spec_from_file_location("ref_src_checkpoint",
                                                      os.path.join(ref_loader.REF_ROOT, "src", "checkpoint.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    tree = _fake_flax_tree(vit_init.reference_state_dict(CFG, seed=4, scaled=True), CFG["num_heads"])
    tree["pre_logits/kernel"] = np.random.RandomState(0).randn(64, 64).astype(np.float32)   # a key with no rule
    keys, vals = list(tree), list(tree.values())
    want = ref.convert_jax_pytorch(keys, vals)
    got = ck.convert_jax_pytorch(keys, vals)
    assert list(got) == list(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k


def test_reference_pth_layout_roundtrip(tmp_path):
    ck = _ckpt()
    sd = vit_init.reference_state_dict(CFG, seed=5, scaled=False)

    class Net(torch.nn.Module):          # stands in for a module with the src key set (CPU: no kernels involved)
        def __init__(self):
            super().__init__()
            self.p = torch.nn.ParameterList([torch.nn.Parameter(v.clone()) for v in sd.values()])

        def state_dict(self, *a, **kw):
            return {k: v.detach() for k, v in zip(sd, self.p)}

    net = Net()
    opt = torch.optim.SGD(net.parameters(), lr=0.03, momentum=0.9)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=0.03, total_steps=10)
    path = ck.save_checkpoint(str(tmp_path), 7, net, opt, sched, best=True)
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"epoch", "state_dict", "optimizer", "lr_scheduler"} and raw["epoch"] == 7   # src/train.py:70-75
    assert os.path.exists(os.path.join(str(tmp_path), "best.pth"))
    got = ck.load_checkpoint(path)                                                               # src/eval.py's loader
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)


def _plain_args():
    return SimpleNamespace(dim=CFG["emb_dim"], mlp_dim=CFG["mlp_dim"], n_layers=CFG["num_layers"], n_heads=CFG["num_heads"],
                           norm_eps=1e-5, use_lora=False, use_reslr=False, block_size=1, dynamic_start_layer=2,
                           dynamic_reserve_initials=1, dynamic_active_target=0.4)


def test_plain_resvit_under_the_mapping_computes_the_src_logits():
    ck = _ckpt()
    sd = vit_init.reference_state_dict(CFG, seed=6, scaled=True)
    mapped, unmatched = ck.src_to_resvit(sd)
    assert sorted(unmatched) == ["classifier.bias", "classifier.weight"]        # res-vit/utils.py:228-278 has no rule
    mapped["classifier.weight"], mapped["classifier.bias"] = sd["classifier.weight"], sd["classifier.bias"]
    g = torch.Generator().manual_seed(0)
    img = torch.randn(3, 3, 32, 32, generator=g)
    labels = torch.randint(0, 10, (3,), generator=g)
    want = vit_oracle.vit_logits(img, sd)
    got = resvit_oracle.resvit_forward(mapped, _plain_args(), img, labels, training=False)["logits"]
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    back, extra = ck.resvit_to_src(mapped, CFG["num_heads"])
    assert extra == ["classifier.weight", "classifier.bias"] or sorted(extra) == ["classifier.bias", "classifier.weight"]
    for k, v in back.items():
        assert v.shape == sd[k].shape and torch.equal(v, sd[k]), k
    assert set(back) == set(sd) - {"classifier.weight", "classifier.bias"}


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")
def test_mapping_equals_the_reference_loader(tmp_path):
    ck = _ckpt()
    stubs = {}
    for name in ("pandas", "PIL", "PIL.Image", "swanlab"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                stubs[name] = types.ModuleType(name)
    if "PIL" in stubs:
        stubs["PIL"].Image = stubs.get("PIL.Image", types.ModuleType("PIL.Image"))
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_resvit_utils", os.path.join(ref_loader.REF_ROOT, "res-vit", "utils.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k in stubs:
            sys.modules.pop(k, None)
    sd = vit_init.reference_state_dict(CFG, seed=8, scaled=True)
    mod, _ = ref_loader.load_resvit_model()
    args = mod.ModelArgs(dim=CFG["emb_dim"], mlp_dim=CFG["mlp_dim"], n_layers=CFG["num_layers"], n_heads=CFG["num_heads"],
                         n_kv_heads=CFG["num_heads"], image_size=(32, 32), patch_size=(16, 16), num_classes=10, use_lora=False,
                         use_reslr=False, device="cpu")
    torch.manual_seed(0)
    ref_model = mod.Transformer(args)
    before = {k: v.clone() for k, v in ref_model.state_dict().items()}
    path = str(tmp_path / "src.pth")
    torch.save({"state_dict": sd}, path)
    ref.load_pretrained_with_mapping(ref_model, path, False, SimpleNamespace(summary_dir=str(tmp_path)))
    mapped, _ = ck.src_to_resvit(sd)
    after = ref_model.state_dict()
    for k, v in after.items():
        if k in mapped:
            assert torch.equal(v, mapped[k]), k                      # same tensors land under the same keys
        else:
            assert torch.equal(v, before[k]), k                      # and nothing else is touched
    assert set(mapped) <= set(after)
