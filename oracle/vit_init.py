"""Re-creates the reference constructor's parameters (src/model.py:161-194) WITHOUT the product package:
the same torch RNG call sequence — Conv2d (kaiming-uniform weight, uniform bias), zeros class token,
randn position embedding, per block LayerNorm (ones/zeros), LinearGeneral randn weights x4 with zero
biases, nn.Linear fc1/fc2, final LayerNorm, classifier nn.Linear — so `torch.manual_seed(s)` yields the
reference's state_dict bit for bit (pinned by tests/test_oracle.py against tests/golden)."""
import torch
import torch.nn as nn

from . import vit_oracle

PRESETS = {   # src/config.py:57-104
    "b16": (16, 768, 3072, 12, 12), "b32": (32, 768, 3072, 12, 12), "l16": (16, 1024, 4096, 16, 24),
    "l32": (32, 1024, 4096, 16, 24), "h14": (14, 1280, 5120, 16, 32),
}


def arch_cfg(arch, image=224, num_classes=1000, num_layers=None):
    p, d, m, h, l = PRESETS[arch]
    return dict(image_size=(image, image), patch_size=(p, p), emb_dim=d, mlp_dim=m, num_heads=h,
                num_layers=l if num_layers is None else num_layers, num_classes=num_classes)


def reference_state_dict(cfg, seed, scaled=True):
    torch.manual_seed(seed)
    h, w = cfg["image_size"]
    fh, fw = cfg["patch_size"]
    D, M, H, L = cfg["emb_dim"], cfg["mlp_dim"], cfg["num_heads"], cfg["num_layers"]
    n = (h // fh) * (w // fw)
    dh = D // H
    sd = {}
    conv = nn.Conv2d(3, D, kernel_size=(fh, fw), stride=(fh, fw))
    cls = torch.zeros(1, 1, D)
    pos = torch.randn(1, n + 1, D)
    layers = []
    for i in range(L):
        n1 = nn.LayerNorm(D)
        qw, qb = torch.randn(D, H, dh), torch.zeros(H, dh)
        kw, kb = torch.randn(D, H, dh), torch.zeros(H, dh)
        vw, vb = torch.randn(D, H, dh), torch.zeros(H, dh)
        ow, ob = torch.randn(H, dh, D), torch.zeros(D)
        n2 = nn.LayerNorm(D)
        fc1 = nn.Linear(D, M)
        fc2 = nn.Linear(M, D)
        layers.append((n1, qw, qb, kw, kb, vw, vb, ow, ob, n2, fc1, fc2))
    norm = nn.LayerNorm(D)
    head = nn.Linear(D, cfg["num_classes"])
    sd["cls_token"] = cls
    sd["embedding.weight"], sd["embedding.bias"] = conv.weight.detach(), conv.bias.detach()
    sd["transformer.pos_embedding.pos_embedding"] = pos
    for i, (n1, qw, qb, kw, kb, vw, vb, ow, ob, n2, fc1, fc2) in enumerate(layers):
        pre = "transformer.encoder_layers.%d." % i
        sd[pre + "norm1.weight"], sd[pre + "norm1.bias"] = n1.weight.detach(), n1.bias.detach()
        for name, wt, bs in (("query", qw, qb), ("key", kw, kb), ("value", vw, vb), ("out", ow, ob)):
            sd[pre + "attn.%s.weight" % name], sd[pre + "attn.%s.bias" % name] = wt, bs
        sd[pre + "norm2.weight"], sd[pre + "norm2.bias"] = n2.weight.detach(), n2.bias.detach()
        sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"] = fc1.weight.detach(), fc1.bias.detach()
        sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"] = fc2.weight.detach(), fc2.bias.detach()
    sd["transformer.norm.weight"], sd["transformer.norm.bias"] = norm.weight.detach(), norm.bias.detach()
    sd["classifier.weight"], sd["classifier.bias"] = head.weight.detach(), head.bias.detach()
    sd = {k: v.clone() for k, v in sd.items()}
    return vit_oracle.scaled_init_(sd) if scaled else sd
