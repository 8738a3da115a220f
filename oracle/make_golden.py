"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container:  python oracle/make_golden.py
The vectors pin oracle/vit_oracle.py (tests/test_oracle.py) and, on the GPU box, the CUDA path.

  vit_tiny.pt    full tensors: scaled-init state_dict of a 2-layer D=128 ViT, images, labels, the
                 reference's logits, loss and every parameter gradient (fp32), plus fp64 logits.
  vit_b16_l2.pt  ViT-B/16 geometry (N=197, D=768, H=12) with 2 layers: construction recipe (seed),
                 weight / gradient fingerprints and full logits — weights are re-created on the test
                 side by seeding the same constructor sequence.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, vit_oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fingerprint(t):
    f = t.detach().double().flatten()
    return {"shape": tuple(t.shape), "sum": float(f.sum()), "abs": float(f.abs().sum()),
            "head": f[:16].float().clone(), "norm": float(f.norm())}


def run_reference(ref, cfg, seed, batch, img_seed, dtype=torch.float32):
    torch.manual_seed(seed)
    model = ref.VisionTransformer(**cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    vit_oracle.scaled_init_(sd)
    model.load_state_dict(sd)
    model = model.to(dtype)
    model.train()
    g = torch.Generator().manual_seed(img_seed)
    h, w = cfg["image_size"]
    img = torch.randn(batch, 3, h, w, generator=g)
    labels = torch.randint(0, cfg["num_classes"], (batch,), generator=g)
    logits = model(img.to(dtype))
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return sd, img, labels, logits.detach(), loss.detach(), grads


def main():
    if not ref_loader.available():
        raise SystemExit("reference not found at %s" % ref_loader.REF_ROOT)
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_src_model()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    tiny = dict(image_size=(32, 32), patch_size=(16, 16), emb_dim=128, mlp_dim=256, num_heads=2, num_layers=2,
                num_classes=10, attn_dropout_rate=0.0, dropout_rate=0.0)
    sd, img, labels, logits, loss, grads = run_reference(ref, tiny, seed=0, batch=3, img_seed=1)
    _, _, _, logits64, loss64, _ = run_reference(ref, tiny, seed=0, batch=3, img_seed=1, dtype=torch.float64)
    torch.save({"cfg": tiny, "seed": 0, "state_dict": sd, "img": img, "labels": labels, "logits": logits,
                "loss": loss, "grads": grads, "logits64": logits64, "loss64": loss64},
               os.path.join(OUT, "vit_tiny.pt"))

    b16 = dict(image_size=(224, 224), patch_size=(16, 16), emb_dim=768, mlp_dim=3072, num_heads=12, num_layers=2,
               num_classes=100, attn_dropout_rate=0.0, dropout_rate=0.0)
    sd, img, labels, logits, loss, grads = run_reference(ref, b16, seed=0, batch=2, img_seed=2)
    torch.save({"cfg": b16, "seed": 0, "batch": 2, "img_seed": 2, "labels": labels, "logits": logits, "loss": loss,
                "weights_fp": {k: fingerprint(v) for k, v in sd.items()},
                "grads_fp": {k: fingerprint(v) for k, v in grads.items()},
                "grad_cls_token": grads["cls_token"], "grad_classifier_bias": grads["classifier.bias"],
                "grad_norm1_weight_l0": grads["transformer.encoder_layers.0.norm1.weight"]},
               os.path.join(OUT, "vit_b16_l2.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()


def resvit_golden():
    """Res-ViT vectors from the unmodified reference: LoRA + router + approximators, block_size 2 and 1,
    training (seeded Gumbel draw, recorded) and eval."""
    mod, _ = ref_loader.load_resvit_model()
    out = {}
    for bs in (2, 1):
        args = mod.ModelArgs(dim=128, mlp_dim=128, n_layers=3, n_heads=2, n_kv_heads=2, lora_rank=8,
                             dynamic_active_target=0.4, dynamic_start_layer=1, dynamic_router_hdim=64,
                             dynamic_reserve_initials=1, low_rank_dim=32, block_size=bs, use_lora=True, use_reslr=True,
                             image_size=(64, 64), patch_size=(16, 16), num_classes=10, device="cpu")
        torch.manual_seed(20 + bs)
        model = mod.Transformer(args)
        with torch.no_grad():
            model.pos_embedding.pos_embedding.mul_(0.02)
            for name, p in model.named_parameters():      # make the router actually split the tokens
                if name.endswith("router.out_conv.4.weight"):
                    p.normal_(0, 0.5)
                if name.endswith("router.out_conv.4.bias"):
                    p.zero_()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        g = torch.Generator().manual_seed(5)
        img = torch.randn(4, 3, 64, 64, generator=g)
        labels = torch.randint(0, 10, (4,), generator=g)
        model.train()
        torch.manual_seed(777)
        c, a, d, e, metric = model(img, labels)
        (1.0 * c + 2.0 * a + 0.5 * d + 0.1 * e).backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        train = dict(c=c.detach(), a=a.detach(), d=d.detach(), e=e.detach(), metric=float(metric['non_low_rank_ratio']),
                     logits=model.logits.detach().clone(), acts=torch.cat(model.acts, -1).detach().clone(), grads=grads,
                     trainable=sorted(k for k, p in model.named_parameters() if p.requires_grad))
        model.eval()
        with torch.no_grad():
            c, a, d, e, metric = model(img, labels)
        ev = dict(c=c.detach(), e=e.detach(), metric=float(metric['non_low_rank_ratio']), logits=model.logits.detach().clone(),
                  acts=torch.cat(model.acts, -1).detach().clone())
        out["bs%d" % bs] = dict(args={k: getattr(args, k) for k in args.__dataclass_fields__}, state_dict=sd, img=img,
                                labels=labels, gumbel_seed=777, train=train, eval=ev)
    torch.save(out, os.path.join(OUT, "resvit_tiny.pt"))
    print("resvit_tiny.pt", os.path.getsize(os.path.join(OUT, "resvit_tiny.pt")))


if __name__ == "__main__" and ref_loader.available():
    resvit_golden()
