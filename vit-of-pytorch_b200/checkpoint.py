"""Checkpoint and weight formats of the reference, so its callers keep working either side of the hot path
(SURVEY.md §8(f) row N2).  Host-side only: no device code, plain torch / numpy.

  load_checkpoint(path)            src/checkpoint.py:7-17 — `.pth` ({'state_dict': ...}) or a Google JAX ViT `.npz`
                                   (read with numpy: the reference needs tensorflow's gfile only to open a local file)
  convert_jax_pytorch(keys, vals)  src/checkpoint.py:39-112 — flax parameter tree -> src/model.py state-dict keys / layouts
  save_checkpoint(...)             src/train.py:69-81 — {'epoch', 'state_dict', 'optimizer', 'lr_scheduler'}, so
                                   src/eval.py (:23-32) reads what this framework trains
  load_state_dict_for_finetune     src/train.py:113-121 — drop the classifier when the class count differs
  src_to_resvit / resvit_to_src    res-vit/utils.py:228-324 — key renaming and the LinearGeneral [D,H,dh] <-> nn.Linear [out,in]
                                   re-layout between the two model families
  load_pretrained_with_mapping     res-vit/utils.py:158-443 — the Res-ViT fine-tune entry: src checkpoint -> Res-ViT model
"""
import os

import numpy as np
import torch

# flax module names -> src/model.py attribute names (src/checkpoint.py:39-77)
_JAX_FIXED = {
    "Transformer": ["transformer"],
    "encoder_norm": ["norm"],
    "kernel": ["weight"],
    "scale": ["weight"],
    "bias": ["bias"],
    "posembed_input": ["pos_embedding"],
    "pos_embedding": ["pos_embedding"],
    "embedding": ["embedding"],
    "head": ["classifier"],
    "cls": ["cls_token"],
}


def _jax_name(part):
    if part in _JAX_FIXED:
        return _JAX_FIXED[part]
    if "encoderblock" in part:
        return ["encoder_layers", part.split("_")[-1]]
    if "LayerNorm" in part:                       # LayerNorm_0 -> norm1, LayerNorm_2 -> norm2 (others vanish)
        idx = part.split("_")[-1]
        return {"0": ["norm1"], "2": ["norm2"]}.get(idx, [])
    if "MlpBlock" in part:
        return ["mlp"]
    if "Dense" in part:                           # Dense_0 -> fc1, Dense_1 -> fc2
        return ["fc%d" % (int(part.split("_")[-1]) + 1)]
    if "MultiHeadDotProductAttention" in part:
        return ["attn"]
    return [part]


def convert_jax_pytorch(keys, values):
    """flax `/`-separated parameter names + arrays -> {src/model.py key: fp32 tensor}.

    Layout rules (src/checkpoint.py:92-110): 2-D kernels are transposed to nn.Linear's [out, in]; the attention
    kernels stay in flax's einsum layout — q/k/v [D, H, dh], out [H, dh, D] — which is exactly LinearGeneral's;
    the patch-embedding conv kernel goes HWIO -> OIHW; everything else is copied."""
    sd = {}
    for key, value in zip(keys, values):
        names = [n for part in key.split("/") for n in _jax_name(part)]
        t = torch.tensor(np.asarray(value), dtype=torch.float32)
        leaf = names[-1] if names else ""
        if t.dim() == 1:
            t = t.squeeze()
        elif t.dim() == 2 and leaf == "weight":
            t = t.T
        elif t.dim() == 4 and leaf == "weight":
            t = t.permute(3, 2, 0, 1)
        sd[".".join(names)] = t
    return sd


def load_jax(path):
    with np.load(path, allow_pickle=False) as z:
        keys = list(z.keys())
        return keys, [z[k] for k in keys]


def load_checkpoint(path, map_location="cpu", trusted=False):
    """state dict from a reference `.pth` or a JAX ViT `.npz` (src/checkpoint.py:7-17).  `.pth` files are unpickled with
    weights_only=True (tensors and plain containers: all the reference's layout holds, src/train.py:69-81);
    trusted=True allows arbitrary pickles, for files you wrote yourself."""
    if path.endswith("npz"):
        return convert_jax_pytorch(*load_jax(path))
    if path.endswith("pth"):
        ck = torch.load(path, map_location=map_location, weights_only=not trusted)
        return ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
    raise ValueError("checkpoint format {} not supported yet!".format(path.split(".")[-1]))


def save_jax_to_pytorch(jax_path, save_dir):
    """src/checkpoint.py:28-33."""
    name = os.path.basename(jax_path).split(".")[0]
    out = os.path.join(save_dir, name + ".pth")
    torch.save({"state_dict": load_checkpoint(jax_path)}, out)
    return out


def save_checkpoint(save_dir, epoch, model, optimizer=None, lr_scheduler=None, best=False):
    """src/train.py:69-81 layout; `model` may be wrapped (DataParallel-style `.module`)."""
    net = model.module if hasattr(model, "module") else model
    state = {"epoch": epoch, "state_dict": {k: v.detach().cpu() for k, v in net.state_dict().items()}}
    if optimizer is not None:
        state["optimizer"] = optimizer.state_dict()
    if lr_scheduler is not None:
        state["lr_scheduler"] = lr_scheduler.state_dict()
    os.makedirs(save_dir, exist_ok=True)
    path = os.path.join(save_dir, "current.pth")
    torch.save(state, path)
    if best:
        torch.save(state, os.path.join(save_dir, "best.pth"))
    return path


def load_state_dict_for_finetune(model, state_dict, num_classes=None):
    """src/train.py:113-121: a head of another width is dropped and re-initialised by the constructor."""
    sd = dict(state_dict)
    if num_classes is None:
        num_classes = model.classifier.weight.shape[0]
    if "classifier.weight" in sd and sd["classifier.weight"].shape[0] != num_classes:
        del sd["classifier.weight"]
        sd.pop("classifier.bias", None)
        return model.load_state_dict(sd, strict=False)
    return model.load_state_dict(sd)


# --------------------------------------------------------------------------------------------------
# src/model.py <-> res-vit/model.py
# --------------------------------------------------------------------------------------------------
_BLOCK_RENAMES = (
    (".attn.query", ".attention.wq"), (".attn.key", ".attention.wk"), (".attn.value", ".attention.wv"),
    (".attn.out", ".attention.wo"), (".mlp.fc1", ".feed_forward.fc1"), (".mlp.fc2", ".feed_forward.fc2"),
    (".norm1", ".attention_norm.layer_norm"), (".norm2", ".ffn_norm.layer_norm"),
)
_TOP_RENAMES = {
    "transformer.norm.weight": "norm.layer_norm.weight", "transformer.norm.bias": "norm.layer_norm.bias",
    "transformer.pos_embedding.pos_embedding": "pos_embedding.pos_embedding",
    "embedding.weight": "embedding.weight", "embedding.bias": "embedding.bias", "cls_token": "cls_token",
}


def map_src_key(key):
    """res-vit/utils.py:228-278; None = no rule (the classifier is deliberately not carried over)."""
    if key.startswith("transformer.encoder_layers."):
        new = key.replace("transformer.encoder_layers.", "layers.")
        for a, b in _BLOCK_RENAMES:
            if a in new:
                return new.replace(a, b)
        return new
    return _TOP_RENAMES.get(key)


def _src_tensor_to_resvit(key, t):
    """LinearGeneral -> nn.Linear layouts (res-vit/utils.py:280-324)."""
    if ".attn.query." in key or ".attn.key." in key or ".attn.value." in key:
        if t.dim() == 3:                                    # weight [D, H, dh] -> [H*dh, D]
            return t.reshape(t.shape[0], -1).transpose(0, 1)
        if t.dim() == 2:                                    # bias [H, dh] -> [H*dh]
            return t.reshape(-1)
    if ".attn.out." in key and t.dim() == 3:                # weight [H, dh, D] -> [D, H*dh]
        return t.reshape(-1, t.shape[2]).transpose(0, 1)
    return t


def src_to_resvit(state_dict):
    """(mapped state dict, unmatched source keys)."""
    out, unmatched = {}, []
    for k, v in state_dict.items():
        nk = map_src_key(k)
        if nk is None:
            unmatched.append(k)
        else:
            out[nk] = _src_tensor_to_resvit(k, v)
    return out, unmatched


def resvit_to_src(state_dict, n_heads):
    """Inverse of src_to_resvit for the backbone tensors (LoRA / router / approximator tensors have no src
    counterpart and are returned as the second value)."""
    inv_top = {v: k for k, v in _TOP_RENAMES.items()}
    out, extra = {}, []
    for k, v in state_dict.items():
        if k in inv_top:
            out[inv_top[k]] = v
            continue
        if not k.startswith("layers."):
            extra.append(k)
            continue
        new = None
        for a, b in _BLOCK_RENAMES:
            if b + "." in k:
                new = k.replace("layers.", "transformer.encoder_layers.", 1).replace(b, a)
                break
        if new is None:
            extra.append(k)
            continue
        t = v
        if ".attn.query." in new or ".attn.key." in new or ".attn.value." in new:
            if t.dim() == 2:                                # [H*dh, D] -> [D, H, dh]
                t = t.transpose(0, 1).reshape(t.shape[1], n_heads, -1)
            else:                                           # [H*dh] -> [H, dh]
                t = t.reshape(n_heads, -1)
        elif ".attn.out.weight" in new:                     # [D, H*dh] -> [H, dh, D]
            t = t.transpose(0, 1).reshape(n_heads, -1, t.shape[0])
        out[new] = t.contiguous()
    return out, extra


def load_pretrained_with_mapping(model, pretrained, strict=False):
    """res-vit/utils.py:158-443 without its JSON side files: loads a src-format checkpoint (path or state dict)
    into a Res-ViT `Transformer`.  Returns (missing_keys, unmatched_keys) like the reference's log summary."""
    sd = load_checkpoint(pretrained) if isinstance(pretrained, str) else pretrained
    target = model.state_dict()
    mapped, unmatched = src_to_resvit(sd)
    new_sd, missing = {}, []
    for k, v in mapped.items():
        if k in target and tuple(v.shape) == tuple(target[k].shape):
            new_sd[k] = v.contiguous()
        elif strict:
            missing.append(k)
        else:
            unmatched.append(k)
    missing += [k for k in target if k not in new_sd]
    model.load_state_dict(new_sd, strict=False)
    return missing, unmatched
