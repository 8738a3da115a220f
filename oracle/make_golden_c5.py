"""Generates tests/golden/resvit_c5.pt and tests/golden/vit_b16_l12.pt by running the UNMODIFIED reference on CPU.

Run in the build container:  python oracle/make_golden_c5.py

  resvit_c5.pt    Res-ViT at the geometry of BASELINE.json configs[4] (D = 768, 197 tokens, 12 heads, router hidden 512,
                  approximator rank 256, LoRA rank 8, target 0.4), 4 layers (dynamic from layer 2), batch 2, block_size 1
                  and 2: construction recipe (seed + the in-place edits below; the product's constructors reproduce the
                  reference state_dict bit for bit, tests/test_resvit_oracle.py), the Gumbel samples the reference drew,
                  losses, logits, router decisions, every trainable gradient as a fingerprint and the small ones in full.
  vit_b16_l12.pt  the full-depth ViT-B/16 (12 layers, 197 tokens, C = 100) forward at batch 8: logits of the fp32
                  reference for the scaled-init recipe — the 2e-2 bf16 bar of the north star at the depth it is quoted for.
"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, resvit_oracle, vit_init, vit_oracle  # noqa: E402
from oracle.make_golden import fingerprint  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FULL_GRAD_LIMIT = 8192     # gradients up to this many elements are stored in full


def c5_args(block_size):
    return dict(dim=768, mlp_dim=3072, n_layers=4, n_heads=12, n_kv_heads=12, lora_rank=8, dynamic_active_target=0.4,
                dynamic_start_layer=2, dynamic_router_hdim=512, dynamic_reserve_initials=1, low_rank_dim=256,
                block_size=block_size, use_lora=True, use_reslr=True, image_size=(224, 224), patch_size=(16, 16),
                num_classes=100, device="cpu")


def c5_edit_(model, seed):
    """In-place edits after construction (same call order on the reference and on the product): trained-like scale of
    the position embedding; a router head that actually splits the tokens; LoRA B and approximator up-projections away
    from their zero init so that every trainable gradient is exercised."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        model.pos_embedding.pos_embedding.mul_(0.02)
        for name, p in model.named_parameters():
            if name.endswith("router.out_conv.4.weight"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.5)
            elif name.endswith("router.out_conv.4.bias"):
                p.zero_()
            elif name.endswith("lora_B.weight") or name.endswith("up_proj.weight"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.02)


def c5_inputs():
    g = torch.Generator().manual_seed(55)
    img = torch.randn(2, 3, 224, 224, generator=g)
    labels = torch.randint(0, 100, (2,), generator=g)
    return img, labels


def resvit_c5():
    mod, _ = ref_loader.load_resvit_model()
    out = {}
    img, labels = c5_inputs()
    for bs in (1, 2):
        kw = c5_args(bs)
        margs = mod.ModelArgs(**kw)
        kw = {k: getattr(margs, k) for k in margs.__dataclass_fields__}      # with the reference's defaults filled in
        torch.manual_seed(40 + bs)
        model = mod.Transformer(margs)
        c5_edit_(model, 400 + bs)
        model.train()
        torch.manual_seed(778)
        c, a, d, e, metric = model(img, labels)
        (1.0 * c + 2.0 * a + 0.5 * d + 0.1 * e).backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        # the Gumbel samples of that run, re-drawn by the oracle under the same seed (and the oracle checked on the way)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        log = []
        torch.manual_seed(778)
        with torch.no_grad():
            o = resvit_oracle.resvit_forward(sd, SimpleNamespace(**kw), img, labels, training=True, noise_log=log)
        acts = torch.cat(model.acts, -1).detach().clone()
        assert torch.equal(o["acts"], acts), "oracle / reference decisions differ"
        assert float((o["logits"] - model.logits).norm() / model.logits.norm()) < 1e-5
        train = dict(c=c.detach(), a=a.detach(), d=d.detach(), e=e.detach(), metric=float(metric["non_low_rank_ratio"]),
                     logits=model.logits.detach().clone(), acts=acts, noise=[t.clone() for t in log],
                     grads_fp={k: fingerprint(v) for k, v in grads.items()},
                     grads={k: v for k, v in grads.items() if v.numel() <= FULL_GRAD_LIMIT},
                     trainable=sorted(k for k, p in model.named_parameters() if p.requires_grad))
        model.eval()
        with torch.no_grad():
            c, a, d, e, metric = model(img, labels)
        ev = dict(c=c.detach(), e=e.detach(), metric=float(metric["non_low_rank_ratio"]), logits=model.logits.detach().clone(),
                  acts=torch.cat(model.acts, -1).detach().clone())
        out["bs%d" % bs] = dict(args=kw, seed=40 + bs, edit_seed=400 + bs, gumbel_seed=778, train=train, eval=ev,
                                weights_fp={k: fingerprint(v) for k, v in sd.items()})
        print("bs", bs, "active", float(metric["non_low_rank_ratio"]), "acts mean", float(acts.float().mean()))
    torch.save(out, os.path.join(OUT, "resvit_c5.pt"))
    print("resvit_c5.pt", os.path.getsize(os.path.join(OUT, "resvit_c5.pt")))


def vit_b16_l12():
    ref = ref_loader.load_src_model()
    cfg = vit_init.arch_cfg("b16", 224, 100)
    sd = vit_init.reference_state_dict(cfg, seed=0, scaled=True)
    model = ref.VisionTransformer(attn_dropout_rate=0.0, dropout_rate=0.0, **cfg)
    model.load_state_dict(sd)
    model.eval()
    g = torch.Generator().manual_seed(12)
    img = torch.randn(8, 3, 224, 224, generator=g)
    with torch.no_grad():
        logits = model(img)
    torch.save({"cfg": cfg, "seed": 0, "img_seed": 12, "batch": 8, "logits": logits.clone()}, os.path.join(OUT, "vit_b16_l12.pt"))
    print("vit_b16_l12.pt", os.path.getsize(os.path.join(OUT, "vit_b16_l12.pt")), "logits std", float(logits.std()))


if __name__ == "__main__":
    if not ref_loader.available():
        raise SystemExit("reference not found at %s" % ref_loader.REF_ROOT)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    resvit_c5()
    vit_b16_l12()
