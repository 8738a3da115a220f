"""torch.autograd.Function layer over the sm_100a kernels (ops.py -> C ABI -> libvitb200.so).

Precision modes (vitb200.set_precision / `with vitb200.precision(...)`):
  "bf16": GEMM/attention operands bf16, fp32 accumulation, fp32 residual stream, fp32 LayerNorm
          statistics, fp32 master weights with bf16 shadows.   (north-star: logits within 2e-2)
  "fp32": parity mode.  Every contraction runs on the SAME tcgen05 kernel as three bf16 segments
          (a = a_hi + a_lo, b = b_hi + b_lo; a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, error ~2^-17), attention
          runs in fp32 on CUDA cores.                           (north-star: rel 1e-4)
Every Function raises on non-CUDA tensors: there is no CPU or eager fallback.
"""
import contextlib
import weakref

import torch

from . import _lib as L
from . import ops

BF16, F32 = torch.bfloat16, torch.float32


class _State:
    precision = "bf16"


def set_precision(p):
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _State.precision = p


def get_precision():
    return _State.precision


@contextlib.contextmanager
def precision(p):
    old = _State.precision
    set_precision(p)
    try:
        yield
    finally:
        _State.precision = old


# --------------------------------------------------------------------------------------------------
# bf16 shadows of the fp32 master weights
# --------------------------------------------------------------------------------------------------
class ShadowStore:
    """bf16 (hi) and residual (lo = w - hi) copies of fp32 weights, refreshed when the master changes.

    A change is detected through the tensor version counter / data pointer; the fused optimizers
    (optim.py) update the masters through raw pointers and write the shadows in the same kernel, so
    for them nothing is re-cast."""

    def __init__(self):
        self._e = {}

    def _entry(self, w):
        key = id(w)
        e = self._e.get(key)
        if e is None or e["ref"]() is not w:
            e = {"ref": weakref.ref(w, lambda _r, k=key: self._e.pop(k, None)), "version": -1, "ptr": 0,
                 "hi": None, "lo": None, "pad": {}}
            self._e[key] = e
        return e

    def get(self, w, want_lo):
        e = self._entry(w)
        stale = (e["version"] != w._version or e["ptr"] != w.data_ptr() or e["hi"] is None
                 or e["hi"].device != w.device or (want_lo and e["lo"] is None))
        if stale:
            src = w.detach()
            if not src.is_contiguous():
                src = src.contiguous()
            if e["hi"] is None or e["hi"].device != w.device or e["hi"].shape != w.shape:
                e["hi"] = torch.empty(w.shape, dtype=BF16, device=w.device)
                e["lo"] = None
            if want_lo and e["lo"] is None:
                e["lo"] = torch.empty(w.shape, dtype=BF16, device=w.device)
            ops.cast_split(src, hi=e["hi"], lo=e["lo"])
            e["version"], e["ptr"] = w._version, w.data_ptr()
            e["pad"] = {}
        return e["hi"], (e["lo"] if want_lo else None)

    def get_padded_2d(self, w, rows, cols, rows_p, cols_p, want_lo):
        """Shadow of w viewed [rows, cols], zero-padded to [rows_p, cols_p] (TMA needs 16-byte row strides,
        i.e. widths that are multiples of 8 bf16 values)."""
        hi, lo = self.get(w, want_lo)
        if rows_p == rows and cols_p == cols:
            return hi.view(rows, cols), (lo.view(rows, cols) if lo is not None else None)
        e = self._entry(w)
        key = (rows_p, cols_p, bool(want_lo))
        if key not in e["pad"]:
            ph = torch.zeros((rows_p, cols_p), dtype=BF16, device=w.device)
            ph[:rows, :cols] = hi.view(rows, cols)
            pl = None
            if want_lo:
                pl = torch.zeros((rows_p, cols_p), dtype=BF16, device=w.device)
                pl[:rows, :cols] = lo.view(rows, cols)
            e["pad"][key] = (ph, pl)
        return e["pad"][key]

    def attach(self, w, hi, lo=None):
        """Use caller-owned shadow buffers (views of a flat optimizer buffer) and fill them now."""
        e = self._entry(w)
        e["hi"], e["lo"] = hi, lo
        ops.cast_split(w.detach().contiguous(), hi=hi, lo=lo)
        e["version"], e["ptr"] = w._version, w.data_ptr()
        e["pad"] = {}

    def mark_fresh(self, w):
        e = self._entry(w)
        e["version"], e["ptr"] = w._version, w.data_ptr()
        e["pad"] = {}


SHADOW = ShadowStore()


# --------------------------------------------------------------------------------------------------
# parameter packs: tensors that one grouped GEMM reads / writes as G matrices at a uniform distance
# --------------------------------------------------------------------------------------------------
def mark_packed(*tensors):
    """Declare that `tensors` (same shape) should be laid out back to back by the fused optimizers' flat buffers, so that
    the merged q|k|v projection finds its three weights / biases / gradients at a uniform stride."""
    group = tuple(tensors)
    for t in group:
        t._vitb_pack = group


def packed_order(params):
    """`params` with the members of every pack moved next to the pack's first member (order otherwise unchanged)."""
    ids = {id(p) for p in params}
    out, placed = [], set()
    for p in params:
        if id(p) in placed:
            continue
        group = getattr(p, "_vitb_pack", None)
        if group is not None and all(id(t) in ids for t in group) and not any(id(t) in placed for t in group):
            for t in group:
                out.append(t)
                placed.add(id(t))
        else:
            out.append(p)
            placed.add(id(p))
    return out


def _uniform_stack(ts):
    """A [G, *shape] strided view over G same-shape contiguous tensors that sit at one distance in ONE storage
    (views of a flat buffer), or None."""
    t0 = ts[0]
    if any(t.shape != t0.shape or t.dtype != t0.dtype or not t.is_contiguous() or t.device != t0.device
           or t.untyped_storage().data_ptr() != t0.untyped_storage().data_ptr() for t in ts):
        return None
    es = t0.element_size()
    step = ts[1].data_ptr() - t0.data_ptr()
    if step < t0.numel() * es or step % es != 0:
        return None
    if any(ts[i].data_ptr() - t0.data_ptr() != i * step for i in range(len(ts))):
        return None
    return torch.as_strided(t0, (len(ts),) + tuple(t0.shape), (step // es,) + tuple(t0.stride()))


class _PackedShadows:
    """bf16 copies of G weights as one [G, ...] tensor for callers without a flat optimizer buffer (inference):
    refreshed when any of the masters changes."""

    def __init__(self):
        self._e = {}

    def get(self, ws):
        key = tuple(id(w) for w in ws)
        sig = tuple((w._version, w.data_ptr()) for w in ws)
        e = self._e.get(key)
        if e is None or e[0] != sig or any(r() is not w for r, w in zip(e[2], ws)) or e[1].device != ws[0].device:
            buf = torch.empty((len(ws),) + tuple(ws[0].shape), dtype=BF16, device=ws[0].device)
            for i, w in enumerate(ws):
                ops.cast_split(w.detach().contiguous(), hi=buf[i])
            e = (sig, buf, tuple(weakref.ref(w, lambda _r, k=key: self._e.pop(k, None)) for w in ws))
            self._e[key] = e
        return e[1]


_PACKED = _PackedShadows()


def packed_weight_operand(ws, shape2d):
    """[G, *shape2d] bf16 operand of the grouped GEMM for the G weights `ws` (bf16 mode): a strided view over the flat
    shadow buffer when the fused optimizer laid them out at one distance, else a cached packed copy."""
    his = [SHADOW.get(w, False)[0] for w in ws]
    st = _uniform_stack(his)
    if st is None:
        st = _PACKED.get(ws)
    return st.view(len(ws), *shape2d)


def _fp32_mode():
    return _State.precision == "fp32"


def _operand(t2d):
    """bf16 GEMM operand pieces of a 2-D activation: [hi] (bf16 mode) or [hi, lo] (fp32 mode)."""
    if t2d.dtype == BF16:
        if t2d.stride(1) != 1:
            t2d = t2d.contiguous()
        return [t2d]
    if t2d.dtype != F32:
        raise L.VitbError("activations must be fp32 or bf16, got %s" % t2d.dtype)
    if not t2d.is_contiguous():
        t2d = t2d.contiguous()
    hi, lo = ops.cast_split(t2d, want_lo=_fp32_mode())
    return [hi, lo] if lo is not None else [hi]


def _weight_operand(w, shape2d):
    hi, lo = SHADOW.get(w, _fp32_mode())
    return [hi.view(shape2d), lo.view(shape2d)] if lo is not None else [hi.view(shape2d)]


def _pairs(a, b):
    """Segment lists for a ~= sum(a) times b ~= sum(b), dropping the lo*lo term."""
    A, B = [a[0]], [b[0]]
    if len(b) > 1:
        A.append(a[0]); B.append(b[1])
    if len(a) > 1:
        A.append(a[1]); B.append(b[0])
    return A, B


def _act_dtype():
    return F32 if _fp32_mode() else BF16


def _grad_target(w):
    """fp32 buffer to accumulate a weight gradient into directly (set by the fused optimizers / DDP)."""
    return getattr(w, "_vitb_main_grad", None)


def _wgrad(dy_ops, x_ops, w, shape2d, dy_is_rows_of_n):
    """dW for a weight viewed `shape2d`.  dy_is_rows_of_n: weight is [N,K] (nn.Linear) -> dW = dY^T X;
    else weight is [K,N] (LinearGeneral) -> dW = X^T dY.  Both operands are consumed MN-major."""
    tgt = _grad_target(w)
    out = tgt.view(shape2d) if tgt is not None else torch.zeros(shape2d, dtype=F32, device=w.device)
    A, B = _pairs(dy_ops, x_ops) if dy_is_rows_of_n else _pairs(x_ops, dy_ops)
    ops.gemm(A, B, a_mn=True, b_mn=True, out=out, accumulate=True)
    return None if tgt is not None else out.view(w.shape)


def _bias_grad(dy2d, b):
    if b is None:
        return None
    tgt = _grad_target(b)
    out = tgt.view(-1) if tgt is not None else torch.zeros(b.numel(), dtype=F32, device=b.device)
    ops.colsum(dy2d, out)
    return None if tgt is not None else out.view(b.shape)


def _as2d(t, cols):
    t2 = t.reshape(-1, cols)
    if t2.stride(1) != 1 or (t2.shape[0] > 1 and t2.stride(0) < cols):
        t2 = t2.contiguous()
    return t2


# --------------------------------------------------------------------------------------------------
# Linear (nn.Linear [N,K] and LinearGeneral [K,N]) with fused bias / GELU / residual epilogue
# --------------------------------------------------------------------------------------------------
def _dy_operand(dy2):
    """bf16 GEMM operand pieces of an incoming fp32 gradient: the bf16 copy its producer left in the backward side
    channel (a LayerNorm backward writes it in the same pass, see _side_put) when there is one, else a cast."""
    if dy2.dtype == F32 and not _fp32_mode() and dy2.is_contiguous():
        side = _side_take(dy2)
        if side is not None:
            return [side[0]]
    return _operand(dy2)


class _Linear(torch.autograd.Function):
    """y = act(x W^T + b + row_bias[row // group]) (+ residual).

    Widths that are not multiples of 8 (the classifier's num_classes, the router's 2*block_size logits) are
    zero-padded to the next multiple of 8 on the weight side, so every TMA row stride stays 16-byte aligned.
    `w_cols=(s, e)` contracts only against columns [s, e) of an nn.Linear weight (the router's split of
    Linear(2H, H) over cat(x_embed, global) into a token GEMM plus a per-image row bias)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, row_bias, layout, act, out_dtype, w_cols, row_bias_group):
        L.require_cuda(x, weight)
        kn = layout == "kn"
        K = x.shape[-1]
        if w_cols is not None:
            if kn:
                raise L.VitbError("linear: w_cols needs an nn.Linear-layout weight")
            N, Kfull = weight.shape
            if w_cols[1] - w_cols[0] != K:
                raise L.VitbError("linear: w_cols span %s does not match the input width %d" % (w_cols, K))
        else:
            N = weight.numel() // K
            Kfull = K
        Np = (N + 7) // 8 * 8
        if Np != N and (kn or residual is not None or act is not None or w_cols is not None or row_bias is not None):
            raise L.VitbError("linear: output width %d must be a multiple of 8 for this layout/epilogue" % N)
        w2 = (K, N) if kn else (N, Kfull)
        x2 = _as2d(x, K)
        xo = _operand(x2)
        if Np != N:
            whi, wlo = SHADOW.get_padded_2d(weight, N, K, Np, K, _fp32_mode())
            wo = [whi] + ([wlo] if wlo is not None else [])
        else:
            wo = _weight_operand(weight, w2)
            if w_cols is not None:
                wo = [w[:, w_cols[0]:w_cols[1]] for w in wo]
        A, B = _pairs(xo, wo)
        M = x2.shape[0]
        out = torch.empty((M, Np), dtype=out_dtype, device=x.device)
        need_z = act == "gelu" and any(ctx.needs_input_grad)
        z = torch.empty((M, N), dtype=out_dtype, device=x.device) if need_z else None
        res2 = _as2d(residual, N) if residual is not None else None
        b1 = bias.detach().view(-1) if bias is not None else None
        if b1 is not None and Np != N:
            b1 = torch.cat([b1, b1.new_zeros(Np - N)])
        if act == "gelu" and residual is not None:
            raise L.VitbError("linear: GELU and residual cannot be fused in one call")
        rb = None
        if row_bias is not None:
            rb = row_bias.detach().reshape(-1, N).float().contiguous()
        ops.gemm(A, B, b_mn=kn, out=out, bias=b1, residual=res2, row_bias=rb, row_bias_group=row_bias_group,
                 epilogue=ops.EPI_GELU if act == "gelu" else ops.EPI_NONE, d2=z)
        ctx.kn, ctx.act, ctx.w2, ctx.N, ctx.K, ctx.Np = kn, act, w2, N, K, Np
        ctx.w_cols, ctx.rb_group = w_cols, row_bias_group
        ctx.rb_shape = row_bias.shape if row_bias is not None else None
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        ctx.has_res = residual is not None
        ctx.res_shape = residual.shape if residual is not None else None
        ctx.res_dtype = residual.dtype if residual is not None else None
        ctx.weight, ctx.bias = weight, bias
        ctx.xo, ctx.z, ctx.wo = xo, z, wo
        if Np != N:
            out = out[:, :N]
        return out.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        N, K, kn, Np = ctx.N, ctx.K, ctx.kn, ctx.Np
        weight, bias = ctx.weight, ctx.bias
        dy2 = _as2d(dy, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dres = None
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = dy.reshape(ctx.res_shape).to(ctx.res_dtype)
        if ctx.act == "gelu":
            dy2 = ops.gelu_bwd(dy2, ctx.z)
        drb = None
        if ctx.rb_shape is not None and ctx.needs_input_grad[4]:
            G = dy2.shape[0] // ctx.rb_group
            drb = (ops.token_mean_fwd(dy2.view(G, ctx.rb_group, N), 0) * float(ctx.rb_group)).reshape(ctx.rb_shape)
        if Np != N:
            dyp = dy2.new_zeros((dy2.shape[0], Np))
            dyp[:, :N] = dy2
            dy2 = dyp
        dyo = _dy_operand(dy2)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            A, B = _pairs(dyo, ctx.wo)
            dx = ops.gemm(A, B, b_mn=not kn, out_dtype=ctx.x_dtype).view(ctx.x_shape)
        if Np == N and ctx.w_cols is None:
            if weight.requires_grad:
                dw = _wgrad(dyo, ctx.xo, weight, ctx.w2, dy_is_rows_of_n=not kn)
            if bias is not None and bias.requires_grad:
                db = _bias_grad(dy2, bias)
        elif ctx.w_cols is not None:
            if weight.requires_grad:
                tgt = _grad_target(weight)
                full = tgt if tgt is not None else torch.zeros(weight.shape, dtype=F32, device=dy.device)
                A, B = _pairs(dyo, ctx.xo)
                ops.gemm(A, B, a_mn=True, b_mn=True, out=full[:, ctx.w_cols[0]:ctx.w_cols[1]], accumulate=True)
                dw = None if tgt is not None else full
            if bias is not None and bias.requires_grad:
                db = _bias_grad(dy2, bias)
        else:
            if weight.requires_grad:
                dwp = torch.zeros((Np, K), dtype=F32, device=dy.device)
                A, B = _pairs(dyo, ctx.xo)
                ops.gemm(A, B, a_mn=True, b_mn=True, out=dwp, accumulate=True)
                dw = dwp[:N].reshape(weight.shape)
            if bias is not None and bias.requires_grad:
                dbp = torch.zeros(Np, dtype=F32, device=dy.device)
                ops.colsum(dy2, dbp)
                db = dbp[:N].reshape(bias.shape)
        ctx.xo = ctx.z = ctx.wo = None
        return dx, dw, db, dres, drb, None, None, None, None, None


def linear(x, weight, bias=None, *, layout="nk", act=None, residual=None, out_dtype=None, w_cols=None,
           row_bias=None, row_bias_group=0):
    """y = act(x W^T + b [+ row_bias]) (+ residual).  layout "nk": weight [N,K] (nn.Linear); "kn": weight
    [K,N...] (LinearGeneral).  Output dtype: fp32 when a residual is fused or in fp32 mode, else bf16."""
    if out_dtype is None:
        out_dtype = F32 if (residual is not None or _fp32_mode()) else BF16
    return _Linear.apply(x, weight, bias, residual, row_bias, layout, act, out_dtype, w_cols, row_bias_group)


# --------------------------------------------------------------------------------------------------
# MLP block: fc2(GELU(fc1(x))) (+ residual), GELU' folded into the fc2 dgrad epilogue
# --------------------------------------------------------------------------------------------------
class _Mlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, out_dtype):
        L.require_cuda(x, w1, w2)
        D = x.shape[-1]
        Mh = w1.shape[0]
        Do = w2.shape[0]
        x2 = _as2d(x, D)
        xo = _operand(x2)
        rows = x2.shape[0]
        adt = _act_dtype()
        h = torch.empty((rows, Mh), dtype=adt, device=x.device)
        need_bwd = any(ctx.needs_input_grad)     # inference: no second output (77 MB per layer at ViT-B/16 batch 128)
        z = torch.empty((rows, Mh), dtype=adt, device=x.device) if need_bwd else None
        A, B = _pairs(xo, _weight_operand(w1, (Mh, D)))
        # bf16 activations: the forward epilogue stores gelu'(z) in place of z (it has exp(-z^2/2) and the cdf at
        # hand anyway), so the fc2 dgrad epilogue is a plain multiply instead of 17 instructions per element
        ctx.z_is_grad = adt == BF16
        ops.gemm(A, B, out=h, bias=b1.detach() if b1 is not None else None,
                 epilogue=ops.EPI_GELU_DG if (ctx.z_is_grad and need_bwd) else ops.EPI_GELU, d2=z)
        ho = _operand(h)
        A, B = _pairs(ho, _weight_operand(w2, (Do, Mh)))
        res2 = _as2d(residual, Do) if residual is not None else None
        out = torch.empty((rows, Do), dtype=out_dtype, device=x.device)
        ops.gemm(A, B, out=out, bias=b2.detach() if b2 is not None else None, residual=res2)
        ctx.dims = (D, Mh, Do)
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        ctx.res = (residual.shape, residual.dtype) if residual is not None else None
        ctx.params = (w1, b1, w2, b2)
        ctx.xo, ctx.ho, ctx.z = xo, ho, z
        return out.view(*x.shape[:-1], Do)

    @staticmethod
    def backward(ctx, dy):
        D, Mh, Do = ctx.dims
        w1, b1, w2, b2 = ctx.params
        dy2 = _as2d(dy, Do)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dres = dy.reshape(ctx.res[0]).to(ctx.res[1]) if (ctx.res is not None and ctx.needs_input_grad[5]) else None
        dyo = _dy_operand(dy2)
        # dz = (dy W2) * gelu'(z)
        A, B = _pairs(dyo, _weight_operand(w2, (Do, Mh)))
        dz = ops.gemm(A, B, b_mn=True, out_dtype=ctx.z.dtype,
                      epilogue=ops.EPI_MUL_AUX if ctx.z_is_grad else ops.EPI_GELU_BWD, aux=ctx.z)
        dzo = _operand(dz)
        dw2 = _wgrad(dyo, ctx.ho, w2, (Do, Mh), True) if w2.requires_grad else None
        db2 = _bias_grad(dy2, b2) if (b2 is not None and b2.requires_grad) else None
        dw1 = _wgrad(dzo, ctx.xo, w1, (Mh, D), True) if w1.requires_grad else None
        db1 = _bias_grad(dz, b1) if (b1 is not None and b1.requires_grad) else None
        dx = None
        if ctx.needs_input_grad[0]:
            A, B = _pairs(dzo, _weight_operand(w1, (Mh, D)))
            dx = ops.gemm(A, B, b_mn=True, out_dtype=ctx.x_dtype).view(ctx.x_shape)
        ctx.xo = ctx.ho = ctx.z = None
        return dx, dw1, db1, dw2, db2, dres, None


def mlp(x, w1, b1, w2, b2, *, residual=None, out_dtype=None):
    if out_dtype is None:
        out_dtype = F32 if (residual is not None or _fp32_mode()) else BF16
    return _Mlp.apply(x, w1, b1, w2, b2, residual, out_dtype)


# --------------------------------------------------------------------------------------------------
# dropout
# --------------------------------------------------------------------------------------------------
class DropoutSite:
    """RNG state of one nn.Dropout of the reference (src/model.py:12,35-36,111): a seed and a draw counter in device
    memory that the kernel itself advances (so a captured step draws a fresh mask per replay).  Not a registered buffer:
    the state-dict keys stay the reference's.  Seeds differ per site and per rank."""
    _sites = 0

    def __init__(self):
        DropoutSite._sites += 1
        self.index = DropoutSite._sites
        self._state = None
        self._seed = None

    def get(self, device):
        if self._state is None or self._state.device != device:
            import torch.distributed as dist
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
            self._seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self.index * 0xD1B54A32D192ED03 + rank * 0x94D049BB133111EB) & ((1 << 64) - 1)
            self._state = torch.zeros(2, dtype=torch.int64, device=device)
        return self._seed, self._state


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, p, site):
        L.require_cuda(x)
        xc = x if x.is_contiguous() else x.contiguous()
        if xc.dtype not in (F32, BF16):
            xc = xc.float()
        res = None
        if residual is not None:
            res = residual if (residual.dtype == F32 and residual.is_contiguous()) else residual.float().contiguous()
        seed, state = site.get(x.device)
        y, mask = ops.dropout_fwd(xc, p, seed, state, residual=res)
        ctx.save_for_backward(mask)
        ctx.p, ctx.x_dtype, ctx.res_dtype = p, x.dtype, (residual.dtype if residual is not None else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        mask, = ctx.saved_tensors
        dx = ops.dropout_bwd(dy, mask, ctx.p, ctx.x_dtype if ctx.x_dtype in (F32, BF16) else F32) if ctx.needs_input_grad[0] else None
        if dx is not None and dx.dtype != ctx.x_dtype:
            dx = dx.to(ctx.x_dtype)
        dres = None
        if ctx.res_dtype is not None and ctx.needs_input_grad[1]:
            dres = dy if dy.dtype == ctx.res_dtype else dy.to(ctx.res_dtype)
        return dx, dres, None, None


def dropout(x, p, training, site, *, residual=None):
    """nn.Dropout(p)(x) [+ residual] (src/model.py:19-20,45-50,122-124).  Identity (plus the residual) when not training
    or p == 0; the mask comes from this library's Philox stream, not torch's generator."""
    if not training or not p:
        return x if residual is None else x.float() + residual
    if p >= 1.0:
        z = torch.zeros_like(x, dtype=F32 if residual is not None else x.dtype)
        return z if residual is None else z + residual
    return _Dropout.apply(x, residual, float(p), site)


# --------------------------------------------------------------------------------------------------
# LayerNorm
# --------------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        L.require_cuda(x, weight, bias)
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        if x2.dtype != F32:
            x2 = x2.float()
        yf, yh, _, mean, rstd = ops.layernorm_fwd(x2, weight.detach(), bias.detach(), eps,
                                                  want_f32=out_dtype == F32, want_bf16=out_dtype == BF16)
        ctx.save_for_backward(x2, mean, rstd)
        ctx.weight, ctx.bias = weight, bias
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        y = yf if out_dtype == F32 else yh
        return y.view(*x.shape[:-1], D)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        D = x2.shape[1]
        dy2 = dy.reshape(-1, D)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        need_wb = weight.requires_grad or bias.requires_grad
        gt, bt = _grad_target(weight), _grad_target(bias)
        dg = (gt if gt is not None else torch.zeros(D, dtype=F32, device=x2.device)) if need_wb else None
        db = (bt if bt is not None else torch.zeros(D, dtype=F32, device=x2.device)) if need_wb else None
        dxf, _, _ = ops.layernorm_bwd(dy2, x2, mean, rstd, weight.detach(), want_f32=True, dgamma=dg, dbeta=db)
        dx = dxf.view(ctx.x_shape)
        if ctx.x_dtype != F32:
            dx = dx.to(ctx.x_dtype)
        return (dx if ctx.needs_input_grad[0] else None,
                dg if (need_wb and gt is None and weight.requires_grad) else None,
                db if (need_wb and bt is None and bias.requires_grad) else None, None, None)


def layer_norm(x, weight, bias, eps=1e-5, *, out_dtype=None):
    """nn.LayerNorm over the last dim.  Output: bf16 GEMM operand in bf16 mode, fp32 in fp32 mode."""
    if out_dtype is None:
        out_dtype = _act_dtype()
    return _LayerNorm.apply(x, weight, bias, eps, out_dtype)


class _LayerNormSkip(torch.autograd.Function):
    """(LayerNorm(x), x) for a pre-LN residual connection, h = x + f(LN(x)): the second output is x itself, to be handed
    to f's last GEMM as its residual.  Forward costs nothing extra; the point is the backward, where the gradient of the
    branch (d LN) and the gradient of the skip arrive at ONE node: the LayerNorm-backward kernel adds the skip gradient in
    its only pass (its `dres` input) and writes the bf16 copy of the sum for the GEMMs upstream — instead of autograd
    adding two fp32 [T, D] tensors in a separate kernel and the upstream GEMM casting the result."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        L.require_cuda(x, weight, bias)
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        yf, yh, _, mean, rstd = ops.layernorm_fwd(x2, weight.detach(), bias.detach(), eps,
                                                  want_f32=out_dtype == F32, want_bf16=out_dtype == BF16)
        ctx.save_for_backward(x2, mean, rstd)
        ctx.weight, ctx.bias = weight, bias
        ctx.x_shape = x.shape
        y = yf if out_dtype == F32 else yh
        return y.view(*x.shape[:-1], D), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        x2, mean, rstd = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        D = x2.shape[1]
        need_wb = weight.requires_grad or bias.requires_grad
        gt, bt = _grad_target(weight), _grad_target(bias)
        dg = (gt if gt is not None else torch.zeros(D, dtype=F32, device=x2.device)) if need_wb else None
        db = (bt if bt is not None else torch.zeros(D, dtype=F32, device=x2.device)) if need_wb else None
        dres = None
        if dskip is not None:
            dres = dskip.reshape(-1, D)
            if dres.dtype != F32 or not dres.is_contiguous():
                dres = dres.float().contiguous()
        if dy is None:       # only the skip was used
            dx = dres.view(ctx.x_shape) if dres is not None else None
        else:
            dy2 = dy.reshape(-1, D)
            if not dy2.is_contiguous():
                dy2 = dy2.contiguous()
            want_bf16 = not _fp32_mode()
            dxf, dxh, _ = ops.layernorm_bwd(dy2, x2, mean, rstd, weight.detach(), dres=dres, want_f32=True,
                                            want_bf16=want_bf16, dgamma=dg, dbeta=db)
            if dxh is not None:
                _side_put(dxf, dxh, None)     # whoever receives exactly this tensor finds its bf16 copy there
            dx = dxf.view(ctx.x_shape)
        return (dx if ctx.needs_input_grad[0] else None,
                dg if (need_wb and gt is None and weight.requires_grad and dy is not None) else None,
                db if (need_wb and bt is None and bias.requires_grad and dy is not None) else None, None, None)


def layer_norm_skip(x, weight, bias, eps=1e-5, *, out_dtype=None):
    """(LayerNorm(x), x): use the second value as the residual of the branch that consumes the first (see _LayerNormSkip).
    Falls back to (layer_norm(x), x) when x is not a contiguous fp32 tensor."""
    if out_dtype is None:
        out_dtype = _act_dtype()
    if x.dtype != F32 or not x.is_contiguous():
        return _LayerNorm.apply(x, weight, bias, eps, out_dtype), x
    return _LayerNormSkip.apply(x, weight, bias, eps, out_dtype)


# --------------------------------------------------------------------------------------------------
# q/k/v projection into one packed [.., 3*HD] buffer, and attention on it
# --------------------------------------------------------------------------------------------------
class _QKVProj(torch.autograd.Function):
    """qkv[..., i*HD:(i+1)*HD] = x W_i (+ b_i) (+ (x A_i^T) B_i^T for LoRA), i in (q, k, v)."""

    @staticmethod
    def forward(ctx, x, layout, wq, bq, wk, bk, wv, bv, *lora):
        L.require_cuda(x, wq, wk, wv)
        kn = layout == "kn"
        K = x.shape[-1]
        ws, bs = (wq, wk, wv), (bq, bk, bv)
        HD = wq.numel() // K
        w2 = (K, HD) if kn else (HD, K)
        x2 = _as2d(x, K)
        xo = _operand(x2)
        rows = x2.shape[0]
        adt = _act_dtype()
        qkv = torch.empty((rows, 3 * HD), dtype=adt, device=x.device)
        has_lora = len(lora) == 6 and lora[0] is not None
        ts = []
        merged = not has_lora and not _fp32_mode() and len(xo) == 1
        if merged:
            _qkv_forward(xo[0], ws, bs, K, HD, qkv, kn=kn)
        for i in range(0 if not merged else 3, 3):
            A, B = _pairs(xo, _weight_operand(ws[i], w2))
            out = qkv[:, i * HD:(i + 1) * HD]
            bias = bs[i].detach().view(-1) if bs[i] is not None else None
            if has_lora:
                la, lb = lora[2 * i], lora[2 * i + 1]          # A [r,K], B [HD,r]  (nn.Linear layouts)
                r = la.shape[0]
                Ar, Br = _pairs(xo, _weight_operand(la, (r, K)))
                t = ops.gemm(Ar, Br, out_dtype=adt)           # t = x A^T  [rows, r]
                to = _operand(t)
                ts.append((t, to))
                lbo = _weight_operand(lb, (HD, r))
                if _fp32_mode():
                    ops.gemm(A, B, b_mn=kn, out=out, bias=bias)
                    A2, B2 = _pairs(to, lbo)
                    ops.gemm(A2, B2, out=out, accumulate=True)
                else:
                    # rank-r update accumulated in the SAME TMEM accumulator as the base projection
                    if kn:
                        raise L.VitbError("LoRA needs nn.Linear-layout base weights")
                    ops.gemm(A + [to[0]], B + [lbo[0]], out=out, bias=bias)
            else:
                ops.gemm(A, B, b_mn=kn, out=out, bias=bias)
        ctx.kn, ctx.K, ctx.HD, ctx.w2 = kn, K, HD, w2
        ctx.ws, ctx.bs, ctx.lora, ctx.has_lora = ws, bs, lora, has_lora
        ctx.xo, ctx.ts = xo, ts
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        return qkv.view(*x.shape[:-1], 3 * HD)

    @staticmethod
    def backward(ctx, dqkv):
        kn, K, HD, w2 = ctx.kn, ctx.K, ctx.HD, ctx.w2
        ws, bs, lora = ctx.ws, ctx.bs, ctx.lora
        d2 = dqkv.reshape(-1, 3 * HD)
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        rows = d2.shape[0]
        fp32 = _fp32_mode()
        if d2.dtype == BF16:
            dyo = [[d2[:, i * HD:(i + 1) * HD]] for i in range(3)]
        else:
            hi, lo = ops.cast_split(d2, want_lo=fp32)
            dyo = [[hi[:, i * HD:(i + 1) * HD]] + ([lo[:, i * HD:(i + 1) * HD]] if lo is not None else [])
                   for i in range(3)]
        grads_w, grads_b, grads_l = [None] * 3, [None] * 3, [None] * 6
        need_dx = ctx.needs_input_grad[0]
        dx = None
        if need_dx:
            if not fp32 and not ctx.has_lora:
                # one GEMM, three K-segments: dX = dQ Wq + dK Wk + dV Wv
                A = [dyo[i][0] for i in range(3)]
                B = [_weight_operand(ws[i], w2)[0] for i in range(3)]
                dx = ops.gemm(A, B, b_mn=not kn, out_dtype=ctx.x_dtype)
            else:
                dx = torch.zeros((rows, K), dtype=F32, device=dqkv.device)
                for i in range(3):
                    A, B = _pairs(dyo[i], _weight_operand(ws[i], w2))
                    ops.gemm(A, B, b_mn=not kn, out=dx, accumulate=True)
        for i in range(3):
            if ws[i].requires_grad:
                grads_w[i] = _wgrad(dyo[i], ctx.xo, ws[i], w2, dy_is_rows_of_n=not kn)
            if bs[i] is not None and bs[i].requires_grad:
                grads_b[i] = _bias_grad(d2[:, i * HD:(i + 1) * HD], bs[i])
            if ctx.has_lora:
                la, lb = lora[2 * i], lora[2 * i + 1]
                r = la.shape[0]
                t, to = ctx.ts[i]
                # dB = dY^T t ; dt = dY B ; dA = dt^T x ; dx += dt A
                if lb.requires_grad:
                    grads_l[2 * i + 1] = _wgrad(dyo[i], to, lb, (HD, r), True)
                A, B = _pairs(dyo[i], _weight_operand(lb, (HD, r)))
                dt = ops.gemm(A, B, b_mn=True, out_dtype=t.dtype)
                dto = _operand(dt)
                if la.requires_grad:
                    grads_l[2 * i] = _wgrad(dto, ctx.xo, la, (r, K), True)
                if need_dx:
                    A, B = _pairs(dto, _weight_operand(la, (r, K)))
                    ops.gemm(A, B, b_mn=True, out=dx, accumulate=True)
        if need_dx:
            if dx.dtype != ctx.x_dtype:
                dx = dx.to(ctx.x_dtype)
            dx = dx.view(ctx.x_shape)
        ctx.xo = ctx.ts = None
        return (dx, None, grads_w[0], grads_b[0], grads_w[1], grads_b[1], grads_w[2], grads_b[2], *grads_l[:len(lora)])


def qkv_proj(x, wq, bq, wk, bk, wv, bv, *, layout, lora=None):
    extra = ()
    if lora is not None:
        extra = tuple(lora)  # (Aq, Bq, Ak, Bk, Av, Bv)
    return _QKVProj.apply(x, layout, wq, bq, wk, bk, wv, bv, *extra)


class _AttentionPacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, H):
        L.require_cuda(qkv)
        B, N, HD3 = qkv.shape
        HD = HD3 // 3
        if not qkv.is_contiguous():
            qkv = qkv.contiguous()
        q, k, v = qkv[:, :, :HD], qkv[:, :, HD:2 * HD], qkv[:, :, 2 * HD:]
        o, lse = ops.attn_fwd(q, k, v, H)
        ctx.save_for_backward(qkv, o, lse)
        ctx.H = H
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        H = ctx.H
        B, N, HD3 = qkv.shape
        HD = HD3 // 3
        if not do.is_contiguous():
            do = do.contiguous()
        q, k, v = qkv[:, :, :HD], qkv[:, :, HD:2 * HD], qkv[:, :, 2 * HD:]
        use_tc = ops.attn_bwd_supported_any(HD // H, N, N, qkv.dtype)
        gdt = BF16 if use_tc else F32
        dqkv = torch.empty((B, N, HD3), dtype=gdt, device=qkv.device)
        ops.attn_bwd(do, q, k, v, o, lse, H, use_tc=use_tc, dq=dqkv[:, :, :HD], dk=dqkv[:, :, HD:2 * HD],
                     dv=dqkv[:, :, 2 * HD:])
        if dqkv.dtype != qkv.dtype:
            dqkv = dqkv.to(qkv.dtype)
        return dqkv, None


def attention_packed(qkv, H):
    """softmax(q k^T / sqrt(dh)) v on a packed [B, N, 3*H*dh] projection; returns [B, N, H*dh]."""
    return _AttentionPacked.apply(qkv, H)


class _AttentionSeparate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, H):
        L.require_cuda(q, k, v)
        q, k, v = (t if t.stride(-1) == 1 else t.contiguous() for t in (q, k, v))
        o, lse = ops.attn_fwd(q, k, v, H)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.H = H
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        if not do.is_contiguous():
            do = do.contiguous()
        dq, dk, dv = ops.attn_bwd(do, q, k, v, o, lse, ctx.H)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None


def attention(q, k, v, H):
    """q [B,Nq,H*dh], k/v [B,Nk,H*dh] -> [B,Nq,H*dh] (asymmetric query/key counts allowed)."""
    return _AttentionSeparate.apply(q, k, v, H)


# --------------------------------------------------------------------------------------------------
# patch embedding: Conv2d(3,D,P,P) + cls token + position embedding as one GEMM with a scatter epilogue
# --------------------------------------------------------------------------------------------------
class PatchColumns:
    """The patch-embedding GEMM operand of a batch, produced without an fp32 image by
    input_pipeline.DeviceImageTransform.patch_columns: hi (and lo in fp32 mode) are bf16 [B*gh*gw, ldk] with rows
    (b, py, px) and k = (c, ph, pw) — exactly what vitb_im2col makes of the loader's fp32 batch."""

    def __init__(self, hi, lo, batch, gh, gw, patch, channels):
        self.hi, self.lo = hi, lo
        self.batch, self.gh, self.gw, self.patch, self.channels = batch, gh, gw, patch, channels

    @property
    def shape(self):
        """Shape of the image batch these columns were cut from: [B, C, gh*P, gw*P]."""
        return (self.batch, self.channels, self.gh * self.patch, self.gw * self.patch)

    @property
    def device(self):
        return self.hi.device

    def to(self, device):
        if torch.device(device) != self.hi.device:
            raise L.VitbError("PatchColumns live on %s; re-create them on %s" % (self.hi.device, device))
        return self


class _PatchEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, conv_w, conv_b, cls_token, pos):
        D, Cin, P, P2 = conv_w.shape
        if P != P2:
            raise L.VitbError("patch embedding needs square patches")
        K = Cin * P * P
        fp32 = _fp32_mode()
        if isinstance(img, PatchColumns):
            L.require_cuda(img.hi, conv_w)
            if img.patch != P or img.channels != Cin:
                raise L.VitbError("PatchColumns were cut for patch %d x %d channels, the embedding is %d x %d"
                                  % (img.patch, img.channels, P, Cin))
            if fp32 and img.lo is None:
                raise L.VitbError("fp32 mode needs PatchColumns made in fp32 mode (low half missing)")
            Bsz, gh, gw = img.batch, img.gh, img.gw
            hi, lo = img.hi, (img.lo if fp32 else None)
        else:
            L.require_cuda(img, conv_w)
            Bsz = img.shape[0]
            img = img.contiguous().float()
            gh, gw = img.shape[2] // P, img.shape[3] // P
            hi, lo = ops.im2col(img, P, want_lo=fp32)
        npatch = gh * gw
        N = npatch + 1
        ldk = hi.shape[1]
        whi, wlo = SHADOW.get_padded_2d(conv_w, D, K, D, ldk, fp32)
        cols = [hi] + ([lo] if lo is not None else [])
        A, Bw = _pairs(cols, [whi] + ([wlo] if wlo is not None else []))
        x = torch.empty((Bsz, N, D), dtype=F32, device=hi.device)
        pos2 = pos.detach().reshape(-1, D)[:N].contiguous() if pos is not None else None
        ops.gemm(A, Bw, out=x.view(Bsz * N, D), bias=conv_b.detach() if conv_b is not None else None,
                 residual=pos2, row_remap_group=npatch)
        ops.cls_rows(x, cls_token.detach().reshape(-1), pos2)
        ctx.cols = cols
        ctx.params = (conv_w, conv_b, cls_token, pos)
        ctx.geom = (Bsz, N, D, K, ldk)
        return x

    @staticmethod
    def backward(ctx, dx):
        conv_w, conv_b, cls_token, pos = ctx.params
        Bsz, N, D, K, ldk = ctx.geom
        dx = dx.contiguous().float()
        _side_take(dx.reshape(-1, D))      # block 0 left its bf16 copy / column sums for a consumer that does not exist
        fp32 = _fp32_mode()
        dev = dx.device
        need_w = conv_w.requires_grad
        dpos = torch.zeros((N, D), dtype=F32, device=dev) if (pos is not None and pos.requires_grad) else None
        dcls = torch.zeros(D, dtype=F32, device=dev) if cls_token.requires_grad else None
        dbias = torch.zeros(D, dtype=F32, device=dev) if (conv_b is not None and conv_b.requires_grad) else None
        phi, plo = ops.embed_bwd(dx, dpos=dpos, dcls=dcls, dbias=dbias, want_patch=need_w, want_lo=fp32)
        dw = None
        if need_w:
            dyo = [phi] + ([plo] if plo is not None else [])
            A, Bm = _pairs(dyo, ctx.cols)
            dwp = torch.zeros((D, ldk), dtype=F32, device=dev)
            ops.gemm(A, Bm, a_mn=True, b_mn=True, out=dwp, accumulate=True)
            dw = dwp[:, :K].reshape(conv_w.shape)
        if dpos is not None:
            full = torch.zeros(pos.shape, dtype=F32, device=dev)
            full.view(-1, D)[:N] = dpos
            dpos = full
        ctx.cols = None
        return (None, dw, dbias, dcls.view(cls_token.shape) if dcls is not None else None, dpos)


def patch_embed(img, conv_w, conv_b, cls_token, pos):
    """[B,3,H,W] -> [B, 1+np, D] fp32 = cat(cls, conv(img)) + pos   (src/model.py:197-204,17)."""
    return _PatchEmbed.apply(img, conv_w, conv_b, cls_token, pos)


# --------------------------------------------------------------------------------------------------
# cross entropy
# --------------------------------------------------------------------------------------------------
class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        L.require_cuda(logits, labels)
        lg = logits.float().contiguous()
        loss, dl = ops.cross_entropy(lg, labels, want_grad=logits.requires_grad)
        ctx.save_for_backward(dl)
        ctx.dtype = logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return (dl * g).to(ctx.dtype), None


def cross_entropy(logits, labels):
    """nn.CrossEntropyLoss() (mean reduction) on [B,C] logits and int64 labels."""
    return _CrossEntropy.apply(logits, labels)


# --------------------------------------------------------------------------------------------------
# fused pre-LN encoder block (bf16 mode): the whole block forward / backward as one autograd node
# --------------------------------------------------------------------------------------------------
# Side channel between consecutive blocks in the backward pass: the block that produces dx (fp32, for autograd) also
# has it in bf16 (the next GEMM operand) and knows its column sums (the bias gradient of the Linear that fed this block's
# input).  An entry is keyed by the storage address of dx and stays valid only while (a) the dx tensor is alive — the
# entry holds a weak reference and disappears with it, so a recycled address can never hit a stale entry — and (b) nobody
# has written into it since (autograd accumulates a second consumer's gradient IN PLACE into a uniquely held buffer,
# which bumps the version counter): the consumer then falls back to casting / summing the tensor it was actually given.
_GRAD_SIDE = {}
SIDE_STATS = [0, 0]     # [entries consumed, entries rejected or missing] — observable by the tests


def _side_put(dx, dxb, colsum):
    import os
    if os.environ.get("VITB_GRAD_SIDE", "1") == "0":
        return
    key = (dx.data_ptr(), dx.numel())
    _GRAD_SIDE[key] = (weakref.ref(dx, lambda _r, k=key: _GRAD_SIDE.pop(k, None)), dx._version, dxb, colsum)


def _side_take(dy2):
    e = _GRAD_SIDE.pop((dy2.data_ptr(), dy2.numel()), None)
    src = e[0]() if e is not None else None
    if src is None or src._version != e[1] or dy2._version != e[1]:
        SIDE_STATS[1] += 1
        return None
    SIDE_STATS[0] += 1
    return e[2], e[3]


_DIRECT = object()      # side-channel column sums that already sit in the upstream block's bias-gradient accumulator


def _upstream_node(x):
    """The fused-block autograd node that produced x (its ctx), or None: lets a block's LayerNorm backward add the column
    sums of dx straight into the fc2 bias gradient of the block upstream instead of handing over a buffer that costs a
    fill and an add launch per block."""
    fn = getattr(x, "grad_fn", None)
    return fn if isinstance(fn, _EncoderBlockFused._backward_cls) else None


def _upstream_bias_target(ctx):
    import os
    up = getattr(ctx, "up", None)
    if up is None or os.environ.get("VITB_GRAD_SIDE", "1") == "0":
        return None
    tgt = _grad_target(up.params[-1])          # the upstream block's fc2 bias
    return tgt.view(-1) if tgt is not None else None


def _acc(param, shape=None):
    """(fp32 accumulator viewed `shape`, value to return to autograd) for a parameter gradient."""
    tgt = _grad_target(param)
    if tgt is not None:
        return (tgt.view(shape) if shape is not None else tgt), None
    buf = torch.zeros(param.shape, dtype=F32, device=param.device)
    return (buf.view(shape) if shape is not None else buf), buf


def _qkv_merge_enabled():
    import os
    return os.environ.get("VITB_QKV_MERGE", "1") != "0"


def _qkv_forward(xn, ws, bs, K, HD, qkv, kn=True):
    """qkv[:, g*HD:(g+1)*HD] = xn W_g + b_g for ws = (wq, wk, wv) in LinearGeneral ("kn": [K, HD]) or nn.Linear ([HD, K])
    layout: one grouped GEMM [T, K] x [K, 3 HD] (src/model.py:86-88 / res-vit/model.py:265-271 as a single contraction)
    when the tile width divides HD, else three GEMMs.  bf16 mode only (callers check)."""
    w2 = (K, HD) if kn else (HD, K)
    if _qkv_merge_enabled() and HD % 256 == 0:
        Bst = packed_weight_operand(ws, w2)
        bias = None
        if bs[0] is not None:
            flat = [b.detach().reshape(-1) for b in bs]
            st = _uniform_stack(flat)
            bias = st.reshape(-1) if (st is not None and st.stride(0) == HD) else torch.cat(flat)
        ops.gemm(xn, Bst, b_mn=kn, out=qkv, bias=bias)
        return
    for i, (w, b) in enumerate(zip(ws, bs)):
        ops.gemm(xn, SHADOW.get(w, False)[0].view(w2), b_mn=kn, out=qkv[:, i * HD:(i + 1) * HD],
                 bias=b.detach().view(-1) if b is not None else None)


class _EncoderBlockFused(torch.autograd.Function):
    """y = h + MLP(LN2(h)),  h = x + Attn(LN1(x))   (src/model.py:117-130), weights in LinearGeneral
    ("kn") layout for attention and nn.Linear ("nk") layout for the MLP.

    Forward: 2 LayerNorm kernels, 4 tcgen05 GEMMs (q | k | v as ONE grouped GEMM into a packed buffer; out+residual;
    fc1+GELU+GELU'; fc2+residual), 1 tcgen05 attention.  Backward: 8 GEMMs (dgrad fc2 x gelu', dW fc2, dW fc1, dgrad
    fc1, dW out, dgrad out, the three q | k | v weight gradients as ONE grouped GEMM, dgrad qkv as three K-segments;
    bias-gradient column sums ride in their epilogues), 1 attention backward, 1 column sum (query bias), 2 LayerNorm
    backwards that add the residual-branch gradient, emit the bf16 operand copy and the neighbouring bias gradients in
    the same pass.  `up`: the fused node that produced x, if any (see _upstream_bias_target)."""

    @staticmethod
    def forward(ctx, x, H, eps1, eps2, n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2, up=None):
        L.require_cuda(x)
        ctx.up = up
        Bsz, N, D = x.shape
        T = Bsz * N
        Mh = w1.shape[0]
        x2 = x.reshape(T, D)
        if x2.dtype != F32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        dev = x.device
        _, xn, _, mean1, rstd1 = ops.layernorm_fwd(x2, n1w.detach(), n1b.detach(), eps1)
        qkv = torch.empty((T, 3 * D), dtype=BF16, device=dev)
        _qkv_forward(xn, (wq, wk, wv), (bq, bk, bv), D, D, qkv)
        qkv3 = qkv.view(Bsz, N, 3 * D)
        o, lse = ops.attn_fwd(qkv3[:, :, :D], qkv3[:, :, D:2 * D], qkv3[:, :, 2 * D:], H)
        o2 = o.view(T, D)
        h = torch.empty((T, D), dtype=F32, device=dev)
        ops.gemm(o2, SHADOW.get(wo, False)[0].view(D, D), b_mn=True, out=h, bias=bo.detach().view(-1), residual=x2)
        _, hn, _, mean2, rstd2 = ops.layernorm_fwd(h, n2w.detach(), n2b.detach(), eps2)
        a = torch.empty((T, Mh), dtype=BF16, device=dev)
        if any(ctx.needs_input_grad):
            z = torch.empty((T, Mh), dtype=BF16, device=dev)
            # z holds gelu'(fc1 pre-activation), not the pre-activation itself (VITB_EPI_GELU_DG / VITB_EPI_MUL_AUX)
            ops.gemm(hn, SHADOW.get(w1, False)[0], out=a, bias=b1.detach(), epilogue=ops.EPI_GELU_DG, d2=z)
        else:       # inference: no derivative output (77 MB per layer at ViT-B/16 batch 128)
            z = a
            ops.gemm(hn, SHADOW.get(w1, False)[0], out=a, bias=b1.detach(), epilogue=ops.EPI_GELU)
        y = torch.empty((T, D), dtype=F32, device=dev)
        ops.gemm(a, SHADOW.get(w2, False)[0], out=y, bias=b2.detach(), residual=h)
        ctx.save_for_backward(x2, xn, mean1, rstd1, qkv, o, lse, h, hn, mean2, rstd2, z, a)
        ctx.params = (n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2)
        ctx.geom = (Bsz, N, D, Mh, H)
        ctx.x_shape = x.shape
        return y.view(Bsz, N, D)

    @staticmethod
    def backward(ctx, dy):
        x2, xn, mean1, rstd1, qkv, o, lse, h, hn, mean2, rstd2, z, a = ctx.saved_tensors
        n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2 = ctx.params
        Bsz, N, D, Mh, H = ctx.geom
        T = Bsz * N
        dev = dy.device
        dy2 = dy.reshape(T, D)
        if dy2.dtype != F32 or not dy2.is_contiguous():
            dy2 = dy2.float().contiguous()
        side = _side_take(dy2)
        # the downstream block may already have added the column sums of ITS dx to this block's fc2 bias gradient
        # (_upstream_bias_target); it left the bf16 copy of that dx here in case the hand-over is rejected
        direct_dxb = getattr(ctx, "b2_direct", None)
        ctx.b2_direct = None
        if side is not None:
            dyb, dy_colsum = side
        else:
            dyb, dy_colsum = ops.cast_split(dy2)[0], None
        grads = {}

        def ret(p, val):
            grads[id(p)] = val

        # ---- MLP ----
        acc_b2, r = _acc(b2)
        ret(b2, r)
        if dy_colsum is _DIRECT:
            pass                               # the downstream block's LayerNorm backward accumulated it in place
        elif dy_colsum is not None:
            acc_b2.add_(dy_colsum)
        else:
            ops.colsum(dy2, acc_b2)
            if direct_dxb is not None:
                # rejected hand-over (a hook or a second consumer changed the gradient on its way here): the sums above are
                # the true ones, so the share the downstream block added in advance comes back out
                share = torch.zeros(D, dtype=F32, device=dev)
                ops.colsum(direct_dxb, share)
                acc_b2.sub_(share)
        acc_b1, r = _acc(b1)
        ret(b1, r)
        dz = torch.empty((T, Mh), dtype=BF16, device=dev)
        ops.gemm(dyb, SHADOW.get(w2, False)[0], b_mn=True, out=dz, epilogue=ops.EPI_MUL_AUX, aux=z, colsum=acc_b1)
        acc, r = _acc(w2)
        ret(w2, r)
        ops.gemm(dyb, a, a_mn=True, b_mn=True, out=acc, accumulate=True)
        acc, r = _acc(w1)
        ret(w1, r)
        ops.gemm(dz, hn, a_mn=True, b_mn=True, out=acc, accumulate=True)
        dhn = ops.gemm(dz, SHADOW.get(w1, False)[0], b_mn=True, out_dtype=BF16)
        del dz
        # ---- LN2 backward + residual: dh = dy + LN'(dhn); column sums of dh = d(out-proj bias) ----
        acc_g2, rg = _acc(n2w)
        acc_be2, rb = _acc(n2b)
        ret(n2w, rg)
        ret(n2b, rb)
        acc_bo, r = _acc(bo)
        ret(bo, r)
        dh, dhb, _ = ops.layernorm_bwd(dhn, h, mean2, rstd2, n2w.detach(), dres=dy2, want_f32=True, want_bf16=True,
                                       dgamma=acc_g2, dbeta=acc_be2, dcolsum=acc_bo.view(-1))
        # ---- attention ----
        acc, r = _acc(wo, (D, D))
        ret(wo, r)
        ops.gemm(o.view(T, D), dhb, a_mn=True, b_mn=True, out=acc, accumulate=True)     # dWo[K,N] = o^T dh
        accs = []
        for b in (bq, bk, bv):
            accb, r = _acc(b)
            ret(b, r)
            accs.append(accb.view(-1))
        do = ops.gemm(dhb, SHADOW.get(wo, False)[0].view(D, D), b_mn=False, out_dtype=BF16, colsum=accs[2])   # + d(bias v)
        qkv3 = qkv.view(Bsz, N, 3 * D)
        dqkv = torch.empty((Bsz, N, 3 * D), dtype=BF16, device=dev)
        ops.attn_bwd(do.view(Bsz, N, D), qkv3[:, :, :D], qkv3[:, :, D:2 * D], qkv3[:, :, 2 * D:], o, lse, H,
                     dq=dqkv[:, :, :D], dk=dqkv[:, :, D:2 * D], dv=dqkv[:, :, 2 * D:])
        dq2 = dqkv.view(T, 3 * D)
        accs_w = []
        for w in (wq, wk, wv):
            acc, r = _acc(w, (D, D))
            ret(w, r)
            accs_w.append(acc)
        gstack = _uniform_stack(accs_w) if _qkv_merge_enabled() and D % 256 == 0 else None
        if gstack is not None:      # the three weight gradients as ONE GEMM: dW[3][K,N] = xn^T [dQ | dK | dV]
            ops.gemm(xn, dq2, a_mn=True, b_mn=True, out=gstack, accumulate=True)
        else:
            for i, acc in enumerate(accs_w):
                ops.gemm(xn, dq2[:, i * D:(i + 1) * D], a_mn=True, b_mn=True, out=acc, accumulate=True)   # dW[K,N] = xn^T dQ
        # bias gradients of q | k | v = column sums of dQ | dK | dV over all tokens.  Two of the three need no pass over
        # dqkv: sum_j dV[j] = sum_i (sum_j P[i,j]) dO[i] = sum_i dO[i] (softmax rows sum to one) — taken by the GEMM that
        # produced dO, in its epilogue — and sum_j dK[j] = sum_i (sum_j dS[i,j]) Q[i] = 0 (sum_j dS[i,j] = D_i - D_i: the
        # scores are invariant to a key bias; the reference's value for it is rounding noise around zero).
        ops.colsum(dq2[:, :D], accs[0])
        dxn = ops.gemm([dq2[:, :D], dq2[:, D:2 * D], dq2[:, 2 * D:]],
                       [SHADOW.get(w, False)[0].view(D, D) for w in (wq, wk, wv)], b_mn=False, out_dtype=BF16)
        # ---- LN1 backward + residual: dx = dh + LN'(dxn) ----
        acc_g1, rg = _acc(n1w)
        acc_be1, rb = _acc(n1b)
        ret(n1w, rg)
        ret(n1b, rb)
        need_dx = ctx.needs_input_grad[0]
        dx = None
        if need_dx:
            direct = _upstream_bias_target(ctx)
            colsum_dx = direct if direct is not None else torch.zeros(D, dtype=F32, device=dev)
            dx, dxb, _ = ops.layernorm_bwd(dxn, x2, mean1, rstd1, n1w.detach(), dres=dh, want_f32=True,
                                           want_bf16=True, dgamma=acc_g1, dbeta=acc_be1, dcolsum=colsum_dx)
            if direct is not None:
                ctx.up.b2_direct = dxb
            _side_put(dx, dxb, _DIRECT if direct is not None else colsum_dx)
            dx = dx.view(ctx.x_shape)
        else:
            ops.layernorm_bwd(dxn, x2, mean1, rstd1, n1w.detach(), dres=dh, want_f32=False, dgamma=acc_g1,
                              dbeta=acc_be1)
        out = [dx, None, None, None]
        for p in ctx.params:
            out.append(grads.get(id(p)) if p.requires_grad else None)
        out.append(None)                       # up
        return tuple(out)


class _EncoderBlockRow0(torch.autograd.Function):
    """Row 0 (the class token) of the LAST encoder block, y0 = h0 + MLP(LN2(h0)), h0 = x0 + Attn(LN1(x))[0]
    (src/model.py:117-130 followed by :155,210, which read nothing but row 0 of the block's output), as one autograd node.

    Keys and values need every token: LayerNorm over all rows, k | v as ONE grouped GEMM into a packed [T, 2D] buffer.
    The query, the output projection and the MLP run on the B class-token rows, read in place through row strides
    (no gather).  Backward: single-query attention backward with bf16 dk | dv straight into the packed buffer, one
    grouped weight-gradient GEMM for k | v, one two-segment dgrad, the query's contribution added to rows 0 of it by a
    GEMM with an in-place residual, and a LayerNorm backward whose residual-branch gradient exists for rows 0 only.
    Every logit and every parameter gradient equals the full block's: the skipped rows have zero gradient there too."""

    @staticmethod
    def forward(ctx, x, H, eps1, eps2, n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2, up=None):
        L.require_cuda(x)
        ctx.up = up
        Bsz, N, D = x.shape
        T = Bsz * N
        Mh = w1.shape[0]
        x2 = x.reshape(T, D)
        if x2.dtype != F32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        dev = x.device
        _, xn, _, mean1, rstd1 = ops.layernorm_fwd(x2, n1w.detach(), n1b.detach(), eps1)
        kv = torch.empty((T, 2 * D), dtype=BF16, device=dev)
        _qkv_forward(xn, (wk, wv), (bk, bv), D, D, kv)
        xn0 = xn.view(Bsz, N * D)[:, :D]            # the class-token rows, in place
        x0 = x2.view(Bsz, N * D)[:, :D]
        q0 = ops.gemm(xn0, SHADOW.get(wq, False)[0].view(D, D), b_mn=True, bias=bq.detach().view(-1), out_dtype=BF16)
        kv3 = kv.view(Bsz, N, 2 * D)
        o0, lse = ops.attn_fwd(q0.view(Bsz, 1, D), kv3[:, :, :D], kv3[:, :, D:], H, use_tc=False)
        h0 = torch.empty((Bsz, D), dtype=F32, device=dev)
        ops.gemm(o0.view(Bsz, D), SHADOW.get(wo, False)[0].view(D, D), b_mn=True, out=h0, bias=bo.detach().view(-1),
                 residual=x0)
        _, hn0, _, mean2, rstd2 = ops.layernorm_fwd(h0, n2w.detach(), n2b.detach(), eps2)
        a0 = torch.empty((Bsz, Mh), dtype=BF16, device=dev)
        if any(ctx.needs_input_grad):
            z0 = torch.empty((Bsz, Mh), dtype=BF16, device=dev)
            ops.gemm(hn0, SHADOW.get(w1, False)[0], out=a0, bias=b1.detach(), epilogue=ops.EPI_GELU_DG, d2=z0)
        else:
            z0 = a0
            ops.gemm(hn0, SHADOW.get(w1, False)[0], out=a0, bias=b1.detach(), epilogue=ops.EPI_GELU)
        y0 = torch.empty((Bsz, D), dtype=F32, device=dev)
        ops.gemm(a0, SHADOW.get(w2, False)[0], out=y0, bias=b2.detach(), residual=h0)
        ctx.save_for_backward(x2, xn, mean1, rstd1, kv, q0, o0, lse, h0, hn0, mean2, rstd2, z0, a0)
        ctx.params = (n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2)
        ctx.geom = (Bsz, N, D, Mh, H)
        ctx.x_shape = x.shape
        return y0

    @staticmethod
    def backward(ctx, dy):
        x2, xn, mean1, rstd1, kv, q0, o0, lse, h0, hn0, mean2, rstd2, z0, a0 = ctx.saved_tensors
        n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2 = ctx.params
        Bsz, N, D, Mh, H = ctx.geom
        T = Bsz * N
        dev = dy.device
        dy2 = dy.reshape(Bsz, D)
        if dy2.dtype != F32 or not dy2.is_contiguous():
            dy2 = dy2.float().contiguous()
        dyb = ops.cast_split(dy2)[0]
        grads = {}

        def acc_of(p, shape=None):
            a, r = _acc(p, shape)
            grads[id(p)] = r
            return a

        # ---- MLP on the B class-token rows ----
        ops.colsum(dy2, acc_of(b2))
        dz = torch.empty((Bsz, Mh), dtype=BF16, device=dev)
        ops.gemm(dyb, SHADOW.get(w2, False)[0], b_mn=True, out=dz, epilogue=ops.EPI_MUL_AUX, aux=z0, colsum=acc_of(b1))
        ops.gemm(dyb, a0, a_mn=True, b_mn=True, out=acc_of(w2), accumulate=True)
        ops.gemm(dz, hn0, a_mn=True, b_mn=True, out=acc_of(w1), accumulate=True)
        dhn = ops.gemm(dz, SHADOW.get(w1, False)[0], b_mn=True, out_dtype=BF16)
        dh0, dhb0, _ = ops.layernorm_bwd(dhn, h0, mean2, rstd2, n2w.detach(), dres=dy2, want_f32=True, want_bf16=True,
                                         dgamma=acc_of(n2w), dbeta=acc_of(n2b), dcolsum=acc_of(bo).view(-1))
        # ---- attention: one query per image ----
        ops.gemm(o0.view(Bsz, D), dhb0, a_mn=True, b_mn=True, out=acc_of(wo, (D, D)), accumulate=True)
        acc_bq, _, acc_bv = [acc_of(b).view(-1) for b in (bq, bk, bv)]      # d(bias k) = 0, see _EncoderBlockFused
        do0 = ops.gemm(dhb0, SHADOW.get(wo, False)[0].view(D, D), b_mn=False, out_dtype=BF16, colsum=acc_bv)
        dkv = torch.empty((T, 2 * D), dtype=BF16, device=dev)
        dkv3 = dkv.view(Bsz, N, 2 * D)
        kv3 = kv.view(Bsz, N, 2 * D)
        dq0 = torch.empty((Bsz, 1, D), dtype=BF16, device=dev)
        ops.attn_q1_bwd(do0.view(Bsz, 1, D), q0.view(Bsz, 1, D), kv3[:, :, :D], kv3[:, :, D:], o0, lse, H,
                        dq=dq0, dk=dkv3[:, :, :D], dv=dkv3[:, :, D:])
        dq2 = dq0.view(Bsz, D)
        xn0 = xn.view(Bsz, N * D)[:, :D]
        ops.gemm(xn0, dq2, a_mn=True, b_mn=True, out=acc_of(wq, (D, D)), accumulate=True)
        ops.colsum(dq2, acc_bq)
        accs_w = [acc_of(w, (D, D)) for w in (wk, wv)]
        gstack = _uniform_stack(accs_w) if _qkv_merge_enabled() and D % 256 == 0 else None
        if gstack is not None:
            ops.gemm(xn, dkv, a_mn=True, b_mn=True, out=gstack, accumulate=True)
        else:
            for i, acc in enumerate(accs_w):
                ops.gemm(xn, dkv[:, i * D:(i + 1) * D], a_mn=True, b_mn=True, out=acc, accumulate=True)
        dx = None
        acc_g1, acc_be1 = acc_of(n1w), acc_of(n1b)
        if ctx.needs_input_grad[0]:
            dxn = ops.gemm([dkv[:, :D], dkv[:, D:]], [SHADOW.get(w, False)[0].view(D, D) for w in (wk, wv)], b_mn=False,
                           out_dtype=BF16)
            dxn0 = dxn.view(Bsz, N * D)[:, :D]
            ops.gemm(dq2, SHADOW.get(wq, False)[0].view(D, D), b_mn=False, out=dxn0, residual=dxn0)   # rows 0 += dq Wq^T
            direct = _upstream_bias_target(ctx)
            colsum_dx = direct if direct is not None else torch.zeros(D, dtype=F32, device=dev)
            dx, dxb, _ = ops.layernorm_bwd(dxn, x2, mean1, rstd1, n1w.detach(), dres=dh0, dres_every=N, want_f32=True,
                                           want_bf16=True, dgamma=acc_g1, dbeta=acc_be1, dcolsum=colsum_dx)
            if direct is not None:
                ctx.up.b2_direct = dxb
            _side_put(dx, dxb, _DIRECT if direct is not None else colsum_dx)
            dx = dx.view(ctx.x_shape)
        else:
            # LN1's own parameters still need the gradient that reaches its output
            dxn = ops.gemm([dkv[:, :D], dkv[:, D:]], [SHADOW.get(w, False)[0].view(D, D) for w in (wk, wv)], b_mn=False,
                           out_dtype=BF16)
            dxn0 = dxn.view(Bsz, N * D)[:, :D]
            ops.gemm(dq2, SHADOW.get(wq, False)[0].view(D, D), b_mn=False, out=dxn0, residual=dxn0)
            ops.layernorm_bwd(dxn, x2, mean1, rstd1, n1w.detach(), want_f32=False, dgamma=acc_g1, dbeta=acc_be1)
        out = [dx, None, None, None]
        for p in ctx.params:
            out.append(grads.get(id(p)) if p.requires_grad else None)
        out.append(None)                       # up
        return tuple(out)


def encoder_block_row0(x, H, norm1, q, k, v, o, norm2, fc1, fc2):
    """Class-token row of the last block as one fused node (bf16 mode, head_dim 64, <= 256 tokens), else None (caller
    composes ops).  VITB_ROW0_FUSED=0 forces the composed path (A/B runs, tests)."""
    import os
    D = x.shape[-1]
    if _fp32_mode() or D % H != 0 or D % 8 != 0 or not ops.attn_q1_supported(D // H, x.shape[1], BF16):
        return None
    if os.environ.get("VITB_ROW0_FUSED", "1") == "0":
        return None
    if any(m.bias is None for m in (q, k, v, o, fc1, fc2)):
        return None
    if not all(p.requires_grad for m in (norm1, q, k, v, o, norm2, fc1, fc2) for p in m.parameters()) and torch.is_grad_enabled():
        return None
    return _EncoderBlockRow0.apply(x, H, norm1.eps, norm2.eps, norm1.weight, norm1.bias, q.weight, q.bias, k.weight,
                                   k.bias, v.weight, v.bias, o.weight, o.bias, norm2.weight, norm2.bias,
                                   fc1.weight, fc1.bias, fc2.weight, fc2.bias, _upstream_node(x))


def encoder_block(x, H, norm1, q, k, v, o, norm2, fc1, fc2):
    """Fused block when possible (bf16 mode, attention shapes the tcgen05 kernels take forward AND backward: head_dim
    64..128 in steps of 16, any token count), else None (caller composes ops)."""
    D = x.shape[-1]
    N = x.shape[1]
    if _fp32_mode() or D % H != 0 or D % 8 != 0 or not (ops.attn_fwd_supported_tc(D // H, N, N, BF16) and
                                                        ops.attn_bwd_supported_any(D // H, N, N, BF16)):
        return None
    if not all(p.requires_grad for m in (norm1, q, k, v, o, norm2, fc1, fc2) for p in m.parameters()) and torch.is_grad_enabled():
        return None
    return _EncoderBlockFused.apply(x, H, norm1.eps, norm2.eps, norm1.weight, norm1.bias, q.weight, q.bias, k.weight,
                                    k.bias, v.weight, v.bias, o.weight, o.bias, norm2.weight, norm2.bias,
                                    fc1.weight, fc1.bias, fc2.weight, fc2.bias, _upstream_node(x))


def clear_grad_side_channel():
    _GRAD_SIDE.clear()


# --------------------------------------------------------------------------------------------------
# Res-ViT routing ops
# --------------------------------------------------------------------------------------------------
class _RouterDecide(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, noise, reserve_initials, training, tau):
        L.require_cuda(logits)
        Bsz, N, bs, two = logits.shape
        lg = logits.float().contiguous()
        nz = noise.float().contiguous() if (training and noise is not None) else None
        soft, hard, ysoft, idx, ent_sum = ops.router_decide_fwd(lg.view(-1, bs, 2), nz, N, bs, reserve_initials, training, tau)
        denom = float(Bsz * (N - reserve_initials) * bs)
        ctx.save_for_backward(soft, ysoft if ysoft is not None else soft)
        ctx.cfg = (Bsz, N, bs, reserve_initials, training, tau, denom, logits.dtype)
        ctx.mark_non_differentiable(idx)
        return hard.view(Bsz, N, bs, 2), idx.view(Bsz, N, 1), ent_sum / denom, soft.view(Bsz, N, bs, 2)

    @staticmethod
    def backward(ctx, d_hard, d_idx, d_ent, d_soft):
        soft, ysoft = ctx.saved_tensors
        Bsz, N, bs, r0, training, tau, denom, dt = ctx.cfg
        dh = d_hard.float().contiguous() if d_hard is not None else None
        ds = d_soft.float().contiguous() if d_soft is not None else None
        de = d_ent.float().contiguous() if d_ent is not None else None
        dl = ops.router_decide_bwd(soft, ysoft, ds, dh, de, 1.0 / denom, N, bs, r0, training, tau)
        return dl.view(Bsz, N, bs, 2).to(dt), None, None, None, None


def router_decide(logits, noise, reserve_initials, training, tau=1.0):
    """-> (hard [B,N,bs,2], indices [B,N,1] fp32 integers, router_entropy scalar, soft [B,N,bs,2])."""
    return _RouterDecide.apply(logits, noise, reserve_initials, training, tau)


class _TokenMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, reserve_initials):
        L.require_cuda(x)
        xc = x.contiguous()
        ctx.cfg = (x.shape[0], x.shape[1], reserve_initials, x.dtype)
        return ops.token_mean_fwd(xc, reserve_initials)

    @staticmethod
    def backward(ctx, dg):
        Bsz, N, r0, dt = ctx.cfg
        return ops.token_mean_bwd(dg.float().contiguous(), Bsz, N, r0, dt), None


def token_mean(x, reserve_initials):
    """[B,N,C] -> [B,C] fp32: mean over the tokens n >= reserve_initials."""
    return _TokenMean.apply(x, reserve_initials)


class _SelectRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, index, member_mask, any_flag=None):
        ref = a if a is not None else b
        L.require_cuda(ref, index)
        cols = ref.shape[-1]
        idx = index.detach().reshape(-1).float().contiguous()
        a2 = a.reshape(-1, cols).contiguous() if a is not None else None
        b2 = b.reshape(-1, cols).contiguous() if b is not None else None
        if a2 is not None and b2 is not None and a2.dtype != b2.dtype:
            a2, b2 = a2.float(), b2.float()
        out = ops.select_rows(a2, b2, idx, member_mask, any_flag)
        ctx.save_for_backward(idx)
        ctx.cfg = (member_mask, a.shape if a is not None else None, b.shape if b is not None else None,
                   a.dtype if a is not None else None, b.dtype if b is not None else None)
        return out.view(ref.shape)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        mask, ash, bsh, adt, bdt = ctx.cfg
        g2 = g.reshape(idx.numel(), -1).contiguous()
        da = db = None
        if ash is not None and ctx.needs_input_grad[0]:
            da = ops.select_rows(g2, None, idx, mask).view(ash).to(adt)
        if bsh is not None and ctx.needs_input_grad[1]:
            db = ops.select_rows(None, g2, idx, mask).view(bsh).to(bdt)
        return da, db, None, None, None


def select_rows(a, b, index, member_ids, any_flag=None):
    """out[t] = a[t] if int(index[t]) in member_ids else b[t]; a or b may be None (zeros).  index: [..., 1]
    fp32 packed router indices; member_ids: iterable of ints < 32 (torch.isin + blend, res-vit/model.py:469-487).
    any_flag: optional int32 CUDA scalar that the kernel sets to 1 when some row is a member."""
    mask = 0
    for i in member_ids:
        mask |= 1 << int(i)
    return _SelectRows.apply(a, b, index, mask, any_flag)


# --------------------------------------------------------------------------------------------------
# Res-ViT scalar losses (one single-CTA kernel each; the kernel also writes the gradient)
# --------------------------------------------------------------------------------------------------
class _DistillLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher):
        L.require_cuda(student, teacher)
        s2 = student if student.dim() == 2 else student.reshape(-1, student.shape[-1])
        t2 = teacher.detach() if teacher.dim() == 2 else teacher.detach().reshape(-1, teacher.shape[-1])
        if s2.dtype != t2.dtype:
            s2, t2 = s2.float(), t2.float()
        if s2.stride(-1) != 1:
            s2 = s2.contiguous()
        if t2.stride(-1) != 1:
            t2 = t2.contiguous()
        loss = torch.zeros((), dtype=F32, device=student.device)
        ds = ops.distill_loss(s2, t2, loss, want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(ds)
        ctx.meta = (student.shape, student.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        (ds,) = ctx.saved_tensors
        shape, dt = ctx.meta
        return (ds * g).view(shape).to(dt), None


def distill_loss(student, teacher):
    """mean((student - teacher.detach())^2)  (DistillLoss, res-vit/model.py:40-59)."""
    return _DistillLoss.apply(student, teacher)


def global_mean_needed(sync_group):
    if sync_group is None or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return False
    return torch.distributed.get_world_size(None if sync_group is True else sync_group) > 1


def global_mean_shift(shard_mean, sync_group):
    """(mean over the process group of `shard_mean`) - shard_mean, as a detached tensor: collective plumbing only (one
    scalar all-reduce; any device / backend), the arithmetic around it is two scalar ops.  Adding it to the shard mean
    gives a quantity whose VALUE is the global-batch mean and whose GRADIENT is this shard's."""
    group = None if sync_group is True else sync_group
    world = torch.distributed.get_world_size(group)
    m = shard_mean.detach().clone()
    torch.distributed.all_reduce(m, op=torch.distributed.ReduceOp.SUM, group=group)
    return (m / world - shard_mean.detach()).contiguous()


class _ActiveLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, reserve_initials, target, sync_group):
        L.require_cuda(probs)
        p3 = probs.float().contiguous()
        shift = None
        if global_mean_needed(sync_group):   # the loss of the GLOBAL batch: one scalar all-reduce of the shard means
            ratio, _, _ = ops.active_loss(p3, reserve_initials, target)
            shift = global_mean_shift(ratio, sync_group)
        _, loss, dp = ops.active_loss(p3, reserve_initials, target, shift=shift, want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(dp)
        ctx.meta = (probs.shape, probs.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dp,) = ctx.saved_tensors
        shape, dt = ctx.meta
        return (dp * g).view(shape).to(dt), None, None, None


def active_loss(probs, reserve_initials, target, sync_group=None):
    """(mean of probs[:, reserve_initials:, :] - target)^2  (ActiveLoss, res-vit/model.py:61-85); with sync_group the mean
    is the global batch's (value) while the gradient stays this shard's — exact once the replicas' gradients are averaged."""
    return _ActiveLoss.apply(probs, reserve_initials, target, sync_group)
