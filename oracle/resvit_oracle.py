"""Functional CPU restatement of the reference's Res-ViT (res-vit/model.py) — the parity oracle for the
router / LoRA / approximator path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pure functions over a state_dict with the reference's key names:
  attention + LoRA      res-vit/model.py:237-299   q = wq(x) + lora_B(lora_A(x)), fp32 softmax, wo
  feed forward          :315-317
  router                :175-211   LN -> Linear -> GELU, global mean over non-reserved tokens, cat, 3 Linears,
                                   softmax, entropy, Gumbel-hard (train) / argmax (eval), reserve override
  bit packing           :169-173
  block forward         :414-529   plain / dynamic-train (teacher, student, blend, approximators) / dynamic-eval
                                   (per-image asymmetric attention loop)
  approximators         :349-368
  model forward         :590-702   losses: CE, ActiveLoss :78-85, DistillLoss :48-59
The Gumbel sample is drawn exactly as torch.nn.functional.gumbel_softmax draws it, in the same order as the
reference (one draw per block-head router, in layer order), so a run seeded like the reference reproduces
its decisions; every drawn sample is appended to `noise_log` so the CUDA path can replay it.
"""
import math

import torch
import torch.nn.functional as TF

from .vit_oracle import gelu, layer_norm, patch_embed


def lra_table(block_size):
    """get_indices_from_LRA_mask (res-vit/model_utils.py:14-107) recomputed from the mapping tables."""
    tables = {
        1: [[[0], []]],
        2: [[[1], [0]], [[], [2]]],
        4: [[[4, 5, 6, 7], [2, 3], [1], [0]], [[], [10, 11], [9], [8]], [[], [], [13, 5], [12, 4]],
            [[], [], [], [2, 6, 10, 14]]],
    }
    tab = tables[block_size]
    out = []
    for j in range(block_size):
        lora = [(i, j) for i in range(j + 1)]
        trans = [(i, jp) for jp in range(j) for i in range(jp + 1)] + \
                [(i, jp) for jp in range(j + 1, block_size) for i in range(j + 1, jp + 1)]
        ste = [(i, jp) for jp in range(j + 1, block_size) for i in range(j + 1)]
        a = sorted({v for (i, jp) in lora for v in tab[i][jp]})
        t = sorted({v for (i, jp) in trans for v in tab[i][jp]} | {(1 << block_size) - 1})
        s = sorted({v for (i, jp) in ste for v in tab[i][jp]})
        out.append((a, t, s))
    return out


def linear(x, sd, pre, bias=True):
    y = x @ sd[pre + ".weight"].t()
    return y + sd[pre + ".bias"] if bias else y


def attention(x, x_kv, sd, pre, n_heads, use_lora):
    squeeze = x.dim() == 2
    if squeeze:
        x, x_kv = x.unsqueeze(0), x_kv.unsqueeze(0)
    B, Nq, D = x.shape
    Nk = x_kv.shape[1]
    dh = D // n_heads

    def proj(t, name):
        y = linear(t, sd, pre + "w" + name)
        if use_lora:
            y = y + (t @ sd[pre + "lora_%s.lora_A.weight" % name].t()) @ sd[pre + "lora_%s.lora_B.weight" % name].t()
        return y

    q = proj(x, "q").view(B, Nq, n_heads, dh).transpose(1, 2)
    k = proj(x_kv, "k").view(B, Nk, n_heads, dh).transpose(1, 2)
    v = proj(x_kv, "v").view(B, Nk, n_heads, dh).transpose(1, 2)
    s = (q @ k.transpose(2, 3)) / math.sqrt(dh)
    p = torch.softmax(s, dim=-1)          # the reference up-casts scores to fp32 (:290); inputs here already are
    o = (p @ v).transpose(1, 2).contiguous().view(B, Nq, -1)
    o = linear(o, sd, pre + "wo")
    return o.squeeze(0) if squeeze else o


def feed_forward(x, sd, pre):
    return linear(gelu(linear(x, sd, pre + "fc1")), sd, pre + "fc2")


def norm(x, sd, pre, eps):
    return layer_norm(x, sd[pre + ".layer_norm.weight"], sd[pre + ".layer_norm.bias"], eps)


def dense_block(x, sd, pre, args):
    xn = norm(x, sd, pre + "attention_norm", args.norm_eps)
    h = x + attention(xn, xn, sd, pre + "attention.", args.n_heads, args.use_lora)
    return h + feed_forward(norm(h, sd, pre + "ffn_norm", args.norm_eps), sd, pre + "feed_forward.")


def router(x, sd, pre, args, training, noise_log):
    B, N, _ = x.shape
    bs, r0 = args.block_size, args.dynamic_reserve_initials
    xe = gelu(linear(norm(x, sd, pre + "in_conv.0", args.norm_eps), sd, pre + "in_conv.1"))
    gfeat = xe[:, r0:, :].mean(dim=1, keepdim=True) if r0 > 0 else xe.mean(dim=1, keepdim=True)
    fused = torch.cat([xe, gfeat.expand(B, N, -1)], dim=-1)
    h = gelu(linear(fused, sd, pre + "out_conv.0"))
    h = gelu(linear(h, sd, pre + "out_conv.2"))
    logits = linear(h, sd, pre + "out_conv.4").view(B, N, bs, 2)
    soft = torch.softmax(logits, dim=-1)
    probs = soft[:, r0:, :, :]
    entropy = -torch.sum(probs * torch.log(probs + 1e-8)) / (B * (N - r0) * bs)
    if training:
        g = -torch.empty_like(logits, memory_format=torch.legacy_contiguous_format).exponential_().log()
        if noise_log is not None:
            noise_log.append(g.detach().clone())
        y = torch.softmax((logits + g) / 1.0, dim=-1)
        idx = y.max(-1, keepdim=True)[1]
        hard = torch.zeros_like(logits).scatter_(-1, idx, 1.0) - y.detach() + y
    else:
        idx = soft.argmax(dim=-1, keepdim=True)
        hard = torch.zeros_like(soft).scatter_(-1, idx, 1.0)
    if r0 > 0:
        hard = hard.clone()
        hard[:, :r0, :, :] = 0
        hard[:, :r0, :, 1] = 1
    keep = hard[:, :, :, 1]
    weights = torch.tensor([2.0 ** (bs - 1 - i) for i in range(bs)], dtype=torch.float32).unsqueeze(-1)
    indices = keep.float() @ weights
    return hard, indices, entropy, soft


def approximators(x, indices, keys, sd, pre, block_size):
    idx = indices.squeeze(-1)
    full = (1 << block_size) - 1
    for key in keys:
        if key == full:
            continue
        m = idx == key
        if m.any():
            x = x.clone()
            sub = x[m]
            x[m] = (sub @ sd[pre + "approximators.%d.down_proj.weight" % key].t()) @ \
                sd[pre + "approximators.%d.up_proj.weight" % key].t() + sub
    return x


def resvit_forward(sd, args, img, labels, training, noise_log=None):
    """Returns a dict with the reference's 5 outputs plus logits / acts / per-block indices."""
    lra = lra_table(args.block_size) if args.use_reslr else None
    emb = patch_embed(img, sd["embedding.weight"], sd["embedding.bias"])
    x = torch.cat([sd["cls_token"].expand(img.shape[0], -1, -1), emb], dim=1)
    pos = sd["pos_embedding.pos_embedding"]
    n = min(x.shape[1], pos.shape[1])
    x = torch.cat([x[:, :n] + pos[:, :n], x[:, n:]], dim=1) if x.shape[1] > pos.shape[1] else x[:, :n] + pos[:, :n]
    B, N, D = x.shape
    acts, soft_probs, indices_by_block = [], [], {}
    d_loss = torch.tensor(0.0)
    r_entropy = torch.tensor(0.0)
    teacher, student = x, x
    info = {}
    for i in range(args.n_layers):
        pre = "layers.%d." % i
        dyn = args.use_reslr and i >= args.dynamic_start_layer
        if not dyn:
            out = dense_block(student, sd, pre, args)
            teacher = student = out
            acts.append(torch.ones(B, N, 1))
            continue
        rel = i - args.dynamic_start_layer
        bid, posb = rel // args.block_size, rel % args.block_size
        if posb == 0:
            hard, indices, ent, soft = router(student, sd, pre + "router.", args, training, noise_log)
            info = dict(routing=hard[:, :, :, 1], indices=indices, soft=soft[:, :, :, 1], head=pre)
            r_entropy = r_entropy + ent
            indices_by_block[bid] = indices.detach().clone()
            if training:
                soft_probs.append(info["soft"])
        w = info["routing"][:, :, posb:posb + 1]
        idx_long = info["indices"].long()
        act_mask = torch.isin(idx_long, torch.tensor(lra[posb][1]))
        apre = info["head"] + "block_path_approximators."
        if training:
            t_out = dense_block(teacher, sd, pre, args)
            full = dense_block(student, sd, pre, args)
            s_out = act_mask * full + (~act_mask) * student
            s_out = approximators(s_out, info["indices"], lra[posb][0], sd, apre, args.block_size)
            d_loss = d_loss + TF.mse_loss(s_out[:, 0, :], t_out[:, 0, :].detach())
            teacher, student = t_out, s_out
        else:
            xn = norm(student, sd, pre + "attention_norm", args.norm_eps)
            am = act_mask.squeeze(-1)
            rows = []
            for b in range(B):
                xq = xn[b:b + 1, am[b], :]
                att = attention(xq, xn[b:b + 1], sd, pre + "attention.", args.n_heads, args.use_lora)
                fa = student[b:b + 1].clone()
                fa[:, am[b], :] = student[b:b + 1, am[b], :] + att
                rows.append(fa)
            h = torch.cat(rows, dim=0)
            out = h + feed_forward(norm(h, sd, pre + "ffn_norm", args.norm_eps), sd, pre + "feed_forward.")
            s_out = act_mask * out + (~act_mask) * student
            student = approximators(s_out, info["indices"], lra[posb][0], sd, apre, args.block_size)
        acts.append(w)
    feat = norm(student, sd, "norm", args.norm_eps)
    logits = linear(feat[:, 0], sd, "classifier")
    c_loss = TF.cross_entropy(logits, labels)
    activation = torch.cat([a.float() for a in acts], dim=-1)
    r0 = args.dynamic_reserve_initials
    if args.use_reslr:
        if soft_probs:
            ratio = torch.cat(soft_probs, dim=-1)[:, r0:, :].mean()
            a_loss = TF.mse_loss(ratio, torch.tensor(args.dynamic_active_target))
        else:
            a_loss = torch.tensor(0.0)
        metric = float(activation[:, r0:, :].mean())
    else:
        a_loss, metric, r_entropy = None, None, torch.tensor(0.0)
    return dict(c_loss=c_loss, a_loss=a_loss, d_loss=d_loss, r_entropy=r_entropy, active_metric=metric, logits=logits,
                acts=activation, indices=indices_by_block)
