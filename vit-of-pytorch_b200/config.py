"""Architecture presets of the reference (src/config.py:57-104; res-vit/config.py:4-46): arch keys
b16 / b32 / l16 / l32 / h14 -> (patch, emb_dim, mlp_dim, heads, layers).  Dropout is 0.0 in every
preset, as in the reference."""

ARCHS = {
    "b16": dict(patch_size=16, emb_dim=768, mlp_dim=3072, num_heads=12, num_layers=12),
    "b32": dict(patch_size=32, emb_dim=768, mlp_dim=3072, num_heads=12, num_layers=12),
    "l16": dict(patch_size=16, emb_dim=1024, mlp_dim=4096, num_heads=16, num_layers=24),
    "l32": dict(patch_size=32, emb_dim=1024, mlp_dim=4096, num_heads=16, num_layers=24),
    "h14": dict(patch_size=14, emb_dim=1280, mlp_dim=5120, num_heads=16, num_layers=32),
}


def get_arch(name):
    if name not in ARCHS:
        raise KeyError("unknown arch %r (choose from %s)" % (name, sorted(ARCHS)))
    cfg = dict(ARCHS[name])
    cfg["attn_dropout_rate"] = 0.0
    cfg["dropout_rate"] = 0.0
    return cfg


def get_b16_config():
    return get_arch("b16")


def get_b32_config():
    return get_arch("b32")


def get_l16_config():
    return get_arch("l16")


def get_l32_config():
    return get_arch("l32")


def get_h14_config():
    return get_arch("h14")


def build_vit(arch, image_size=224, num_classes=1000):
    """VisionTransformer for an arch key, the way src/train.py:103-112 builds it."""
    from .model import VisionTransformer

    c = get_arch(arch)
    return VisionTransformer(image_size=(image_size, image_size), patch_size=(c["patch_size"], c["patch_size"]),
                             emb_dim=c["emb_dim"], mlp_dim=c["mlp_dim"], num_heads=c["num_heads"],
                             num_layers=c["num_layers"], num_classes=num_classes,
                             attn_dropout_rate=c["attn_dropout_rate"], dropout_rate=c["dropout_rate"])
