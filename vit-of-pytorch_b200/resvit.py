"""Drop-in mirror of the reference's Res-ViT module API (res-vit/model.py), backed by the sm_100a kernels.

Class names, constructor signatures, parameter names / shapes / init order and state_dict keys follow
/root/reference/res-vit/model.py: ModelArgs :13, DistillLoss :40, ActiveLoss :61, PositionEmbs :87,
LoRAModule :104, LayerNorm :119, RouterModule :133, Attention :213, FeedForward :302,
LowRankApproximator :320, BlockPathApproximators :336, TransformerBlock :371, Transformer :532.

What runs underneath (functional.py -> libvitb200.so):
  * LoRA: the rank-r update (x A^T) B^T is a second K-segment of the base q/k/v GEMM — it accumulates in the
    same TMEM accumulator before the epilogue (bf16 mode).
  * Router: LN + GEMM(+GELU) chain; the concat with the per-image global feature is split algebraically into
    a token GEMM plus a per-image row bias; softmax / entropy / Gumbel-or-argmax / reserve override / bit
    packing are one kernel; torch.isin + blends are a 32-bit lookup inside the row-select kernel.
  * Approximators: x[m] += up(down(x[m])) becomes x + up(rowmask(down(x))) — dense, no boolean gather, no
    `.item()` / `.any()` host syncs (res-vit/model.py:357,364).
  * Eval-mode asymmetric attention (per-image Python loop, :503-516): query rows are independent, so full
    attention followed by a row select gives the same values for the active rows without the loop.
  * The teacher path never receives gradients in the reference (every use is detached, :57, :630-633), so
    it runs without building an autograd graph; when teacher and student inputs are the same tensor
    (first dynamic layer, :442) the dense block is computed once.
"""
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
from torch.nn import Linear

from . import functional as F
from .lra_tables import get_indices_from_LRA_mask


@dataclass
class ModelArgs:
    dim: int = 768
    mlp_dim: int = 3072
    n_layers: int = 12
    n_heads: int = 12
    n_kv_heads: Optional[int] = 12
    norm_eps: float = 1e-5
    lora_rank: int = 8
    dynamic_active_target: float = 0.4
    dynamic_start_layer: int = 2
    dynamic_router_hdim: int = 512
    dynamic_reserve_initials: int = 1
    low_rank_dim: int = 256
    block_size: int = 2
    use_lora: bool = False
    use_reslr: bool = False
    image_size: Tuple[int, int] = (224, 224)
    patch_size: Tuple[int, int] = (16, 16)
    num_classes: int = 100
    dropout: float = 0.15
    num_patches: int = (image_size[0] // patch_size[0]) * (image_size[1] // patch_size[1])
    device: str = 'cuda'


class DistillLoss(nn.Module):
    """MSE between student and (detached) teacher class tokens (res-vit/model.py:40-59)."""

    def __init__(self):
        super().__init__()
        self.criterion = torch.nn.MSELoss()

    def forward(self, student_cls, teacher_cls):
        return F.distill_loss(student_cls, teacher_cls)      # vitb_distill_loss: value + gradient in one kernel


class ActiveLoss(nn.Module):
    """(mean keep-probability over non-reserved tokens - target)^2 (res-vit/model.py:61-85).

    The loss is not linear in the batch mean, so under data parallelism the average of the per-replica losses is not the
    loss of the global batch.  `sync_group` selects the semantics:
      None (default)  per-replica, i.e. what each process computes on its shard;
      a process group (or True for the default group)  the GLOBAL batch: one scalar all-reduce of the shard means, and a
      surrogate whose value is (m - t)^2 with m the global mean and whose gradient, once the data-parallel wrapper has
      averaged the replicas' gradients, is exactly d/dtheta (m - t)^2 — equal shard sizes assumed, as in the wrapper.
    """

    def __init__(self, target, reserve_initials, sync_group=None):
        super().__init__()
        self.target = target
        self.reserve_initials = reserve_initials
        self.sync_group = sync_group

    @torch.no_grad()
    def metric(self, activation):
        activation = activation[:, self.reserve_initials:, :]
        return {'non_low_rank_ratio': activation.float().mean(), 'current_target': self.target}

    def forward(self, activation):
        # vitb_active_loss: masked mean, (m - t)^2 and the gradient in one kernel (+ one scalar all-reduce when synced)
        return F.active_loss(activation, self.reserve_initials, self.target, self.sync_group)


class PositionEmbs(nn.Module):
    """x + pos (truncating on a length mismatch, res-vit/model.py:87-101).  Transformer.forward folds the add
    into the patch-embedding GEMM epilogue when the lengths agree."""

    def __init__(self, num_patches, emb_dim):
        super().__init__()
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, emb_dim))

    def forward(self, x):
        n, p = x.shape[1], self.pos_embedding.shape[1]
        if n == p:
            return x + self.pos_embedding
        m = min(n, p)
        out = x[:, :m] + self.pos_embedding[:, :m]
        return torch.cat([out, x[:, m:]], dim=1) if n > p else out


class LoRAModule(nn.Module):
    """lora_B(lora_A(x)), no bias, no scaling, both N(0, 0.01) (res-vit/model.py:104-117)."""

    def __init__(self, in_dim: int, rank: int, out_dim: int):
        super().__init__()
        self.in_dim, self.rank, self.out_dim = in_dim, rank, out_dim
        self.lora_A = nn.Linear(in_dim, rank, bias=False)
        self.lora_B = nn.Linear(rank, out_dim, bias=False)
        nn.init.normal_(self.lora_A.weight, mean=0.0, std=0.01)
        nn.init.normal_(self.lora_B.weight, mean=0.0, std=0.01)

    def forward(self, x, residual=None):
        t = F.linear(x, self.lora_A.weight)
        return F.linear(t, self.lora_B.weight, residual=residual)


class LayerNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6, use_lora: bool = False):
        super().__init__()
        self.layer_norm = nn.LayerNorm(dim, eps=eps)
        if use_lora:
            for p in self.layer_norm.parameters():
                p.requires_grad = False

    def forward(self, x, out_dtype=None):
        ln = self.layer_norm
        return F.layer_norm(x, ln.weight, ln.bias, ln.eps, out_dtype=out_dtype)


class RouterModule(nn.Module):
    """DynamicViT-style 2-way (skip / keep) router per (token, position in block) — res-vit/model.py:133-211."""

    def __init__(self, in_dim: int, hidden_dim: int, reserve_initials: int, norm_eps: float, block_size: int = 1,
                 use_lora: bool = False):
        super().__init__()
        self.block_size = block_size
        self.reserve_initials = reserve_initials
        self.in_conv = nn.Sequential(LayerNorm(in_dim, norm_eps, use_lora=use_lora), nn.Linear(in_dim, hidden_dim),
                                     nn.GELU())
        self.out_conv = nn.Sequential(nn.Linear(hidden_dim * 2, hidden_dim), nn.GELU(),
                                      nn.Linear(hidden_dim, hidden_dim // 2), nn.GELU(),
                                      nn.Linear(hidden_dim // 2, block_size * 2))
        nn.init.normal_(self.out_conv[-1].weight, mean=0, std=0.01)
        for i in range(block_size):
            self.out_conv[-1].bias.data[i * 2] = 0.0
            self.out_conv[-1].bias.data[i * 2 + 1] = 5.0
        self.noise_fn = None   # tests inject the Gumbel sample the oracle drew; default: torch's own draw

    def _router2indices(self, x):
        n = x.shape[-1]
        w = torch.tensor([2.0 ** (n - 1 - i) for i in range(n)], dtype=torch.float32, device=x.device)
        return (x.float() * w).sum(-1, keepdim=True)

    def _gumbel(self, logits):
        if self.noise_fn is not None:
            return self.noise_fn(logits)
        # exactly torch.nn.functional.gumbel_softmax's sample
        return -torch.empty_like(logits, memory_format=torch.legacy_contiguous_format).exponential_().log()

    def forward(self, x):
        B, N, C = x.shape
        hd = self.in_conv[1].out_features
        r0 = self.reserve_initials
        xn = self.in_conv[0](x)
        x_embed = F.linear(xn, self.in_conv[1].weight, self.in_conv[1].bias, act="gelu")          # [B,N,hd]
        g = F.token_mean(x_embed, r0 if r0 > 0 else 0)                                            # [B,hd] fp32
        l0 = self.out_conv[0]
        # Linear(2*hd, hd)(cat(x_embed, g)) = x_embed W[:, :hd]^T + (g W[:, hd:]^T + b)
        row_bias = F.linear(g, l0.weight, l0.bias, w_cols=(hd, 2 * hd), out_dtype=torch.float32)   # [B,hd]
        h1 = F.linear(x_embed, l0.weight, None, w_cols=(0, hd), row_bias=row_bias, row_bias_group=N, act="gelu")
        h2 = F.linear(h1, self.out_conv[2].weight, self.out_conv[2].bias, act="gelu")
        logits = F.linear(h2, self.out_conv[4].weight, self.out_conv[4].bias, out_dtype=torch.float32)
        logits = logits.reshape(B, N, self.block_size, 2)
        noise = self._gumbel(logits) if self.training else None
        hard, indices, entropy, soft = F.router_decide(logits, noise, r0, self.training)
        return hard, indices, entropy, soft


class Attention(nn.Module):
    """Multi-head attention with optional LoRA on q/k/v and optional asymmetric key/value source
    (res-vit/model.py:213-299)."""

    def __init__(self, args: ModelArgs):
        super().__init__()
        self.n_kv_heads = args.n_heads if args.n_kv_heads is None else args.n_kv_heads
        self.n_local_heads = args.n_heads
        self.n_local_kv_heads = self.n_kv_heads
        self.n_rep = self.n_local_heads // self.n_local_kv_heads
        self.head_dim = args.dim // args.n_heads
        self.use_lora = args.use_lora
        self.wq = Linear(args.dim, args.n_heads * self.head_dim, bias=True)
        self.wk = Linear(args.dim, self.n_kv_heads * self.head_dim, bias=True)
        self.wv = Linear(args.dim, self.n_kv_heads * self.head_dim, bias=True)
        self.wo = Linear(args.n_heads * self.head_dim, args.dim, bias=True)
        if self.use_lora:
            self.lora_q = LoRAModule(args.dim, args.lora_rank, self.head_dim * self.n_local_heads)
            self.lora_k = LoRAModule(args.dim, args.lora_rank, self.head_dim * self.n_local_kv_heads)
            self.lora_v = LoRAModule(args.dim, args.lora_rank, self.head_dim * self.n_local_kv_heads)

    def _proj(self, x, lin, lora):
        y = F.linear(x, lin.weight, lin.bias)
        return lora(x, residual=y) if lora is not None else y

    def forward(self, x, x_kv=None, residual=None):
        no_batch = x.dim() == 2
        if no_batch:
            x = x.unsqueeze(0)
            x_kv = x_kv.unsqueeze(0) if x_kv is not None else None
            residual = residual.unsqueeze(0) if residual is not None else None
        if self.n_rep != 1:
            raise NotImplementedError("grouped-query attention (n_kv_heads < n_heads) is not implemented")
        if x_kv is None or x_kv is x:
            lora = None
            if self.use_lora:
                lora = (self.lora_q.lora_A.weight, self.lora_q.lora_B.weight, self.lora_k.lora_A.weight,
                        self.lora_k.lora_B.weight, self.lora_v.lora_A.weight, self.lora_v.lora_B.weight)
            qkv = F.qkv_proj(x, self.wq.weight, self.wq.bias, self.wk.weight, self.wk.bias, self.wv.weight,
                             self.wv.bias, layout="nk", lora=lora)
            o = F.attention_packed(qkv, self.n_local_heads)
        else:
            q = self._proj(x, self.wq, self.lora_q if self.use_lora else None)
            k = self._proj(x_kv, self.wk, self.lora_k if self.use_lora else None)
            v = self._proj(x_kv, self.wv, self.lora_v if self.use_lora else None)
            o = F.attention(q, k, v, self.n_local_heads)
        out = F.linear(o, self.wo.weight, self.wo.bias, residual=residual)
        return out.squeeze(0) if no_batch else out


class FeedForward(nn.Module):
    def __init__(self, dim: int, mlp_dim: int):
        super().__init__()
        self.fc1 = Linear(dim, mlp_dim, bias=True)
        self.fc2 = Linear(mlp_dim, dim, bias=True)
        self.act = nn.GELU()

    def forward(self, x, residual=None):
        return F.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual=residual)


class LowRankApproximator(nn.Module):
    def __init__(self, dim: int, rank: int):
        super().__init__()
        self.down_proj = nn.Linear(dim, rank, bias=False)
        self.up_proj = nn.Linear(rank, dim, bias=False)
        nn.init.normal_(self.down_proj.weight, mean=0.0, std=0.01)
        nn.init.normal_(self.up_proj.weight, mean=0.0, std=0.01)

    def forward(self, x, residual=None):
        return F.linear(F.linear(x, self.down_proj.weight), self.up_proj.weight, residual=residual)


class BlockPathApproximators(nn.Module):
    """x[idx == key] += up_key(down_key(x[idx == key])) for every key of this layer (res-vit/model.py:336-368).

    up_proj has no bias, so up(rowmask * down(x)) == rowmask * up(down(x)): the mask is applied to the small
    [T, rank] tensor and the add rides in the up-projection's residual epilogue.  Tokens carry exactly one
    packed index, so the keys' row sets are disjoint and their order does not matter."""

    def __init__(self, dim: int, rank: int, block_size: int):
        super().__init__()
        self.block_size = block_size
        self.approximators = nn.ModuleDict()
        total = 2 ** block_size
        for key in range(total):
            if key == total - 1:
                continue
            self.approximators[str(key)] = LowRankApproximator(dim, rank)
        self._live_flags = {}       # key -> int32 device flag "this approximator saw a token this step" (bind_optimizer)

    def forward(self, x, router_indices, LRA_mask):
        keys = LRA_mask.tolist() if torch.is_tensor(LRA_mask) else list(LRA_mask)
        for key in keys:
            ks = str(int(key))
            if ks not in self.approximators:
                continue
            ap = self.approximators[ks]
            t = F.linear(x, ap.down_proj.weight)
            # the reference runs the approximator only `if sub_mask.any()` (res-vit/model.py:363-367): whether it ran —
            # whether its parameters have a gradient at all — is recorded for the optimizer by the selection kernel
            t = F.select_rows(t, None, router_indices, [int(key)], any_flag=self._live_flags.get(ks) if self.training else None)
            x = F.linear(t, ap.up_proj.weight, residual=x)
        return x


def bind_optimizer(model, optimizer):
    """Tell a FusedAdamW which parameters may go without a gradient in a step: the members of every
    BlockPathApproximators, whose modules the reference only runs for keys that occur in the batch
    (res-vit/model.py:363-367) — torch.optim.AdamW then skips them entirely (`p.grad is None`: no weight decay, no moment
    decay, no step).  Without this call such parameters see a zero gradient and are decayed like the rest."""
    n = 0
    for m in model.modules():
        if isinstance(m, BlockPathApproximators):
            for ks, ap in m.approximators.items():
                ps = [p for p in ap.parameters() if p.requires_grad]
                if ps:
                    m._live_flags[ks] = optimizer.register_skippable(ps)
                    n += 1
    return n


def _compaction_enabled():
    """VITB_RESVIT_COMPACT (read per call): 1 (default) = inference skips the inactive tokens' output projection and MLP
    (device-side row compaction); 0 = the dense-select path (every row computed, rows selected afterwards)."""
    import os
    return os.environ.get("VITB_RESVIT_COMPACT", "1") != "0"


class TransformerBlock(nn.Module):
    def __init__(self, layer_id: int, args: ModelArgs):
        super().__init__()
        self.n_heads = args.n_heads
        self.dim = args.dim
        self.head_dim = args.dim // args.n_heads
        self.layer_id = layer_id
        self.current_epoch = 0
        self.use_lora = args.use_lora
        self.use_reslr = args.use_reslr
        self.attention = Attention(args)
        self.attention_norm = LayerNorm(args.dim, eps=args.norm_eps, use_lora=args.use_lora)
        self.feed_forward = FeedForward(dim=args.dim, mlp_dim=args.mlp_dim)
        self.ffn_norm = LayerNorm(args.dim, eps=args.norm_eps, use_lora=args.use_lora)
        self.dynamic_start_layer = args.dynamic_start_layer
        if self.use_reslr and self.layer_id >= args.dynamic_start_layer:
            self.block_size = args.block_size
            rel = self.layer_id - self.dynamic_start_layer
            self.is_block_head = rel % self.block_size == 0
            self.current_block_id = rel // self.block_size
            self.block_start_layer = self.dynamic_start_layer + self.current_block_id * self.block_size
            self.current_block_pos = self.layer_id - self.block_start_layer
            if self.is_block_head:
                self.router = RouterModule(args.dim, args.dynamic_router_hdim, args.dynamic_reserve_initials,
                                           args.norm_eps, block_size=self.block_size, use_lora=args.use_lora)
                self.block_path_approximators = BlockPathApproximators(args.dim, args.low_rank_dim, self.block_size)

    def _dense(self, x):
        """h = x + attn(LN(x)); out = h + ffn(LN(h)) — both adds in the producing GEMM's epilogue."""
        x = x if x.dtype == torch.float32 else x.float()
        if torch.is_grad_enabled() and x.requires_grad:
            # LayerNorm and the skip connection as one autograd node: its backward adds the two gradients in the LayerNorm
            # kernel and leaves the bf16 copy for the GEMMs upstream (functional._LayerNormSkip)
            an, fn = self.attention_norm.layer_norm, self.ffn_norm.layer_norm
            xn, xr = F.layer_norm_skip(x, an.weight, an.bias, an.eps)
            h = self.attention(xn, residual=xr)
            hn, hr = F.layer_norm_skip(h, fn.weight, fn.bias, fn.eps)
            return h, self.feed_forward(hn, residual=hr)
        h = self.attention(self.attention_norm(x), residual=x)
        return h, self.feed_forward(self.ffn_norm(h), residual=h)

    def _eval_compacted(self, x, router_indices, transformer_ids):
        """Inference with real token skipping (res-vit/model.py:503-524): the attention output projection and the whole
        MLP run on the ACTIVE rows only, compacted on the device.  Attention itself stays dense over the queries — output
        row i depends on q_i alone, and at 4 % of the layer's FLOPs a variable-length query kernel would buy little —
        so q | k | v and softmax(qk^T)v see every token; then
            rows, count = active rows (device list + device count, nothing is read back)
            h_c  = o[rows] Wo^T + bo + x[rows]          GEMM over `count` rows (vitb_gemm_params.m_dev)
            y_c  = h_c + fc2(GELU(fc1(LN(h_c))))        same
            out  = x, out[rows] = y_c
        which is exactly `mask * output + (~mask) * x` of the reference: the skipped rows' attention / MLP results are
        never used there either.  Buffers keep the capacity of the dense tensors (B * N rows)."""
        from . import ops
        att, ff = self.attention, self.feed_forward
        bsz, seqlen, D = x.shape
        x2 = x.reshape(bsz * seqlen, D)
        xn = self.attention_norm(x)
        lora = None
        if att.use_lora:
            lora = (att.lora_q.lora_A.weight, att.lora_q.lora_B.weight, att.lora_k.lora_A.weight,
                    att.lora_k.lora_B.weight, att.lora_v.lora_A.weight, att.lora_v.lora_B.weight)
        qkv = F.qkv_proj(xn, att.wq.weight, att.wq.bias, att.wk.weight, att.wk.bias, att.wv.weight, att.wv.bias,
                         layout="nk", lora=lora)
        o = F.attention_packed(qkv, att.n_local_heads).reshape(bsz * seqlen, D)
        rows, count = ops.compact_rows(router_indices, transformer_ids)
        o_c = ops.gather_rows(o, rows, count)
        x_c = ops.gather_rows(x2, rows, count)
        h_c = ops.gemm(o_c, F.SHADOW.get(att.wo.weight, False)[0], out_dtype=torch.float32, bias=att.wo.bias.detach(),
                       residual=x_c, m_dev=count)
        hn_c = self.ffn_norm(h_c)                                  # (over the capacity; rows past count are scratch)
        a_c = ops.gemm(hn_c, F.SHADOW.get(ff.fc1.weight, False)[0], bias=ff.fc1.bias.detach(), epilogue=ops.EPI_GELU, m_dev=count)
        y_c = ops.gemm(a_c, F.SHADOW.get(ff.fc2.weight, False)[0], out_dtype=torch.float32, bias=ff.fc2.bias.detach(),
                       residual=h_c, m_dev=count)
        out = x2.clone()
        ops.scatter_rows(y_c, rows, count, out)
        return out.view(bsz, seqlen, D)

    def forward(self, x, teacher_x=None, block_info=None, LRA_mask=None):
        bsz, seqlen, _ = x.shape
        if block_info is None:
            block_info = {}
        if not self.use_reslr or self.layer_id < self.dynamic_start_layer:
            w = torch.ones((bsz, seqlen, 1), device=x.device)
            _, out = self._dense(x)
            return (out, out, w, block_info) if self.training else (out, w, block_info)

        bid = self.current_block_id
        if self.is_block_head:
            routing, router_indices, router_entropy, soft_routing = self.router(x)
            block_info = {
                f"block_{bid}_approximators": self.block_path_approximators,
                f"block_{bid}_routing": routing[:, :, :, 1],
                f"block_{bid}_router_indices": router_indices,
                f"block_{bid}_router_entropy": router_entropy,
                f"block_{bid}_soft_routing": soft_routing[:, :, :, 1],
            }
        approximators = block_info[f"block_{bid}_approximators"]
        block_routing = block_info[f"block_{bid}_routing"]
        router_indices = block_info[f"block_{bid}_router_indices"]
        pos = self.current_block_pos
        w = block_routing[:, :, pos:pos + 1]
        assert LRA_mask is not None, "LRA_mask must be provided"
        approx_ids, transformer_ids = LRA_mask[pos][0], LRA_mask[pos][1]

        if self.training:
            same_input = teacher_x is None or teacher_x is x
            _, transformer_out = self._dense(x)
            if same_input:
                teacher_out = transformer_out.detach()
            else:
                with torch.no_grad():          # no gradient ever reaches the teacher path in the reference
                    _, teacher_out = self._dense(teacher_x)
            student_out = F.select_rows(transformer_out, x, router_indices, transformer_ids)
            student_out = approximators(student_out, router_indices, approx_ids)
            return teacher_out, student_out, w, block_info
        x = x if x.dtype == torch.float32 else x.float()
        if _compaction_enabled() and not torch.is_grad_enabled() and F.get_precision() == "bf16":
            student_out = self._eval_compacted(x, router_indices, transformer_ids)
            student_out = approximators(student_out, router_indices, approx_ids)
            return student_out, w, block_info
        # dense eval: attention for all rows (row i depends on q_i only), keep it on the active rows
        attn_plus_x = self.attention(self.attention_norm(x), residual=x)
        h = F.select_rows(attn_plus_x, x, router_indices, transformer_ids)
        output = self.feed_forward(self.ffn_norm(h), residual=h)
        student_out = F.select_rows(output, x, router_indices, transformer_ids)
        student_out = approximators(student_out, router_indices, approx_ids)
        return student_out, w, block_info


class Transformer(nn.Module):
    def __init__(self, params: ModelArgs):
        super().__init__()
        self.device = params.device
        h, w = params.image_size
        fh, fw = params.patch_size
        params.num_patches = (h // fh) * (w // fw)
        self.embedding = nn.Conv2d(3, params.dim, kernel_size=(fh, fw), stride=(fh, fw))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, params.dim))
        self.pos_embedding = PositionEmbs(params.num_patches, params.dim)
        self.criterion = torch.nn.CrossEntropyLoss()
        self.criterion_active = ActiveLoss(target=params.dynamic_active_target,
                                           reserve_initials=params.dynamic_reserve_initials)
        self.criterion_distill = DistillLoss()
        self.n_layers = params.n_layers
        self.layers = torch.nn.ModuleList()
        for layer_id in range(params.n_layers):
            self.layers.append(TransformerBlock(layer_id, params))
        self.norm = LayerNorm(params.dim, eps=params.norm_eps, use_lora=params.use_lora)
        self.classifier = Linear(params.dim, params.num_classes)
        self.use_lora = params.use_lora
        self.use_reslr = params.use_reslr
        if self.use_lora:
            frozen = ('.feed_forward.', '.attention.wo.', '.attention.wq.', '.attention.wk.', '.attention.wv.')
            for name, p in self.named_parameters():
                if name.startswith('embedding.') or name.startswith('pos_embedding.') or any(f in name for f in frozen):
                    p.requires_grad = False
        if self.use_reslr:
            self.LRA_mask = get_indices_from_LRA_mask(params.block_size)

    def _embed(self, x):
        n = (x.shape[2] // self.embedding.kernel_size[0]) * (x.shape[3] // self.embedding.kernel_size[1]) + 1
        pos = self.pos_embedding.pos_embedding
        if pos.shape[1] == n:
            return F.patch_embed(x, self.embedding.weight, self.embedding.bias, self.cls_token, pos)
        return self.pos_embedding(F.patch_embed(x, self.embedding.weight, self.embedding.bias, self.cls_token, None))

    def forward(self, x, labels):
        device = next(self.parameters()).device
        x = x.to(device) if x.device != device else x
        labels = labels.to(device) if labels.device != device else labels
        x = self._embed(x)
        self.acts, self.soft_routing_probs, self.routing_maps = [], [], {}
        d_loss = torch.zeros((), device=x.device)      # kernel fills (no host->device copy: graph-capturable)
        r_entropy = torch.zeros((), device=x.device)
        block_info = {}
        teacher_x = student_x = x
        for layer in self.layers:
            dynamic = self.use_reslr and layer.layer_id >= layer.dynamic_start_layer
            if self.training:
                teacher_out, student_out, w, block_info = layer(student_x, teacher_x, block_info,
                                                                self.LRA_mask if dynamic else None)
                if dynamic:
                    d_loss = d_loss + self.criterion_distill(student_out[:, 0, :], teacher_out[:, 0, :])
                teacher_x, student_x = teacher_out, student_out
            else:
                student_x, w, block_info = layer(student_x, None, block_info, self.LRA_mask if dynamic else None)
            if dynamic and layer.is_block_head:
                bid = layer.current_block_id
                r_entropy = r_entropy + block_info[f"block_{bid}_router_entropy"]
                self.routing_maps[bid] = block_info[f"block_{bid}_routing"].detach()
                if self.training:
                    self.soft_routing_probs.append(block_info[f"block_{bid}_soft_routing"])
            self.acts.append(w)
        cls = self.norm(student_x[:, 0], out_dtype=torch.float32)   # the head only reads row 0 (:679)
        activation = torch.cat([a.float() for a in self.acts], dim=-1)
        output = F.linear(cls, self.classifier.weight, self.classifier.bias, out_dtype=torch.float32)
        self.logits = output
        c_loss = F.cross_entropy(output, labels)
        if self.use_reslr:
            if len(self.soft_routing_probs) > 0:
                a_loss = self.criterion_active(torch.cat(self.soft_routing_probs, dim=-1))
            else:
                a_loss = torch.zeros((), device=x.device)
            active_metric = self.criterion_active.metric(activation)
        else:
            a_loss, active_metric = None, None
            r_entropy = torch.zeros((), device=x.device)
        return c_loss, a_loss, d_loss, r_entropy, active_metric
