"""Fused optimizers over flat buffers: one kernel per step updates every parameter, its momentum and
the bf16 GEMM shadow of the weights (vitb_sgd_momentum / vitb_adamw in the C ABI).

  FusedSGD   == torch.optim.SGD(lr, momentum, dampening, weight_decay, nesterov)   src/train.py:154-158
  FusedAdamW == torch.optim.AdamW(lr, betas, eps, weight_decay) + clip_grad_norm_  res-vit/train.py:65,272-277

Both re-home the parameters into ONE contiguous fp32 buffer, give every parameter a persistent
gradient view into ONE contiguous fp32 gradient buffer (the autograd Functions accumulate their weight
gradients straight into it — no per-tensor .grad allocation, one memset per step, and the data-parallel
wrapper all-reduces slices of the same buffer), and attach views of ONE bf16 buffer as weight shadows.
They subclass torch.optim.Optimizer so LR schedulers (OneCycleLR, cosine) drive `param_groups[...]['lr']`.
"""
import torch

from . import functional as F
from . import ops

_ALIGN = 64  # elements: 256-byte aligned fp32 slices, 128-byte aligned bf16 slices (TMA needs 16)


class _FlatGroup:
    def __init__(self, params, grad_buffer_factory=None):
        params = F.packed_order([p for p in params if p.requires_grad])
        if not params:
            raise ValueError("no trainable parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("fused optimizers need the model on a CUDA (B200) device — call .cuda() first")
        offs, total = [], 0
        for p in params:
            if p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("fused optimizers need fp32 parameters on one device")
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.params, self.offsets, self.total = params, offs, total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        # the gradient buffer may come from outside: p2p.NvlinkExchange.allocate puts it into symmetric memory
        self.flat_g = (grad_buffer_factory(total, dev) if grad_buffer_factory is not None
                       else torch.zeros(total, dtype=torch.float32, device=dev))
        if self.flat_g.numel() != total or self.flat_g.dtype != torch.float32 or not self.flat_g.is_contiguous():
            raise ValueError("grad_buffer_factory must return a contiguous fp32 tensor of the requested size")
        self.flat_hi = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.flat_lo = None
        for p, off in zip(params, offs):
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off:off + n].view(p.shape)
            gview = self.flat_g[off:off + n].view(p.shape)
            p.grad = gview
            p._vitb_main_grad = gview
            F.SHADOW.attach(p, self.flat_hi[off:off + n].view(p.shape))

    def ensure_lo(self):
        if self.flat_lo is None:
            self.flat_lo = torch.zeros(self.total, dtype=torch.bfloat16, device=self.flat_p.device)
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                F.SHADOW.attach(p, self.flat_hi[off:off + n].view(p.shape), self.flat_lo[off:off + n].view(p.shape))

    def zero_grad(self):
        self.flat_g.zero_()
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
                p._vitb_main_grad = p.grad

    def mark_fresh(self):
        for p in self.params:
            F.SHADOW.mark_fresh(p)


class _FusedBase(torch.optim.Optimizer):
    def _init_flat(self, grad_buffer_factory=None):
        if grad_buffer_factory is not None and len(self.param_groups) != 1:
            raise NotImplementedError("grad_buffer_factory expects one parameter group")
        self._flat = [_FlatGroup(g["params"], grad_buffer_factory) for g in self.param_groups]
        # device-resident hyper-parameters (4 floats per group, see _hyper): a captured CUDA graph of step() keeps
        # following whatever an LR scheduler writes into param_groups (OneCycleLR cycles lr AND momentum / beta1)
        self._hyper_host = [self._hyper(g) for g in self.param_groups]
        self._hyper_dev = [torch.tensor(h, dtype=torch.float32, device=fg.flat_p.device)
                           for h, fg in zip(self._hyper_host, self._flat)]

    def _hyper(self, group):          # -> 4 floats, the layout the kernel expects
        raise NotImplementedError

    def sync_lr(self):
        """Push param_groups[...] (lr, momentum / betas, weight decay) to the device scalars; called by step() outside
        stream capture and by GraphedTrainStep before every replay."""
        for i, g in enumerate(self.param_groups):
            h, old = self._hyper(g), self._hyper_host[i]
            if h != old:
                for j in range(4):      # asynchronous element fills (a host -> device copy would sync the stream)
                    if old is None or h[j] != old[j]:
                        self._hyper_dev[i][j].fill_(h[j])
                self._hyper_host[i] = h

    sync_hyper = sync_lr

    def zero_grad(self, set_to_none=False):  # gradients live in the flat buffer; never set to None
        F.clear_grad_side_channel()
        for fg in self._flat:
            fg.zero_grad()

    def _check_grad_views(self):
        """Weight gradients are accumulated straight into the flat buffer through p._vitb_main_grad; `model.zero_grad()`
        or `p.grad = None` would leave that buffer un-zeroed while hiding it from view.  Refuse to step on such a state."""
        for fg in self._flat:
            base = fg.flat_g.data_ptr()
            for p, off in zip(fg.params, fg.offsets):
                if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                    raise RuntimeError(
                        "fused optimizer: a parameter's .grad no longer aliases the flat gradient buffer — clear gradients "
                        "with optimizer.zero_grad() (not model.zero_grad() / p.grad = None): the kernels accumulate weight "
                        "gradients into the optimizer's flat buffer, which only optimizer.zero_grad() zeroes")

    def flat_grads(self):
        """The contiguous gradient buffers (one per param group) — what the data-parallel wrapper all-reduces."""
        return [fg.flat_g for fg in self._flat]

    def flat_params(self):
        return [fg.flat_p for fg in self._flat]

    # ---- checkpointing: the moments live in flat buffers outside Optimizer.state; export / import them per parameter
    # in torch.optim's own layout so that a resumed run continues exactly (and torch.optim can read the file)
    def _state_buffers(self):         # -> {state key: [flat buffer or None per group]}
        raise NotImplementedError

    def _extra_state(self):
        return {}

    def _load_extra_state(self, extra):
        pass

    def state_dict(self):
        sd = super().state_dict()
        state, idx = {}, 0
        bufs = self._state_buffers()
        for gi, fg in enumerate(self._flat):
            order = {id(p): i for i, p in enumerate(fg.params)}
            for p in self.param_groups[gi]["params"]:
                j = order.get(id(p))
                if j is not None:
                    off, n = fg.offsets[j], p.numel()
                    entry = {k: b[gi][off:off + n].view(p.shape).clone() for k, b in bufs.items() if b[gi] is not None}
                    entry.update({k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._extra_state().items()})
                    if entry:
                        state[idx] = entry
                idx += 1
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for g, saved in zip(self.param_groups, groups):
            if len(saved["params"]) != len(g["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            for k, v in saved.items():
                if k != "params":
                    g[k] = v
        bufs = self._state_buffers()
        idx, extra = 0, None
        for gi, fg in enumerate(self._flat):
            order = {id(p): i for i, p in enumerate(fg.params)}
            for p in self.param_groups[gi]["params"]:
                entry = state_dict["state"].get(idx)
                j = order.get(id(p))
                if entry is not None and j is not None:
                    off, n = fg.offsets[j], p.numel()
                    for k, b in bufs.items():
                        if b[gi] is not None and k in entry:
                            b[gi][off:off + n].copy_(entry[k].reshape(-1).to(b[gi].device, torch.float32))
                    extra = entry
                idx += 1
        if extra is not None:
            self._load_extra_state(extra)
        self._hyper_host = [None] * len(self.param_groups)   # force a refresh of the device scalars
        self.sync_lr()


class FusedSGD(_FusedBase):
    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False,
                 grad_buffer_factory=None):
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        self._init_flat(grad_buffer_factory)
        self._bufs = [torch.zeros_like(fg.flat_p) if g["momentum"] != 0 else None
                      for fg, g in zip(self._flat, self.param_groups)]
        self._steps = 0

    def _hyper(self, g):
        return [float(g["lr"]), float(g["momentum"]), float(g["dampening"]), float(g["weight_decay"])]

    def _state_buffers(self):
        return {"momentum_buffer": self._bufs}

    def _extra_state(self):
        return {"vitb_steps": self._steps}

    def _load_extra_state(self, extra):
        self._steps = int(extra.get("vitb_steps", 1 if "momentum_buffer" in extra else 0))

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedSGD.step does not take a closure")
        want_lo = F.get_precision() == "fp32"
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
            self._check_grad_views()
        for fg, g, buf, hyper in zip(self._flat, self.param_groups, self._bufs, self._hyper_dev):
            if want_lo:
                fg.ensure_lo()
            ops.sgd_momentum(fg.flat_p, fg.flat_g, buf, g["lr"], g["momentum"], dampening=g["dampening"],
                             weight_decay=g["weight_decay"], nesterov=g["nesterov"], first_step=self._steps == 0,
                             shadow_hi=fg.flat_hi, shadow_lo=fg.flat_lo, hyper_dev=hyper)
            fg.mark_fresh()
        self._steps += 1


class FusedAdamW(_FusedBase):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None,
                 grad_buffer_factory=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._init_flat(grad_buffer_factory)
        self._m = [torch.zeros_like(fg.flat_p) for fg in self._flat]
        self._v = [torch.zeros_like(fg.flat_p) for fg in self._flat]
        self.max_grad_norm = max_grad_norm
        dev = self._flat[0].flat_p.device
        self._sumsq = torch.zeros((), dtype=torch.float32, device=dev)
        self._coef = torch.ones((), dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)  # last total norm (device scalar)
        self._steps = 0
        # device-resident step counter: step() can be captured in a CUDA graph
        self._step_dev = torch.zeros((), dtype=torch.int32, device=dev)
        self._seg = None            # per-parameter segment tables, built by the first register_skippable()

    def _hyper(self, g):
        return [float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["weight_decay"])]

    # ---- parameters that may go without a gradient (torch.optim.AdamW skips `p.grad is None`: no weight decay, no moment
    # decay, the parameter's own step count stands still).  With flat buffers "no gradient" has to be said explicitly:
    # register_skippable(params) returns an int32 device flag; whoever computes those parameters' gradients sets it
    # non-zero during the step (Res-ViT: BlockPathApproximators, through vitb_select_rows_flag, when the approximator's
    # key occurred in the batch); zero_grad() clears it.  Unregistered parameters are always updated.
    def register_skippable(self, params):
        params = list(params)
        if self._seg is None:
            self._build_segments()
        k = self._n_flags
        if k >= self._flags.numel():
            raise RuntimeError("FusedAdamW: more than %d skippable parameter sets" % self._flags.numel())
        self._n_flags += 1
        for p in params:
            loc = self._seg_of.get(id(p))
            if loc is None:
                raise ValueError("register_skippable: parameter is not managed by this optimizer")
            gi, j = loc
            self._seg[gi][1][j] = k         # seg_flag (device int32): a tiny element write
        return self._flags[k:k + 1]

    def _build_segments(self):
        dev = self._flat[0].flat_p.device
        self._flags = torch.zeros(256, dtype=torch.int32, device=dev)
        self._n_flags = 0
        self._seg, self._seg_of = [], {}
        for gi, fg in enumerate(self._flat):
            ends = fg.offsets[1:] + [fg.total]
            seg_end = torch.tensor(ends, dtype=torch.int64, device=dev)
            seg_flag = torch.full((len(ends),), -1, dtype=torch.int32, device=dev)
            seg_step = torch.full((len(ends),), float(self._steps), dtype=torch.float32, device=dev)
            self._seg.append((seg_end, seg_flag, seg_step))
            for j, p in enumerate(fg.params):
                self._seg_of[id(p)] = (gi, j)

    def zero_grad(self, set_to_none=False):
        super().zero_grad(set_to_none)
        if self._seg is not None:
            self._flags.zero_()

    def _state_buffers(self):
        return {"exp_avg": self._m, "exp_avg_sq": self._v}

    def _extra_state(self):
        return {"step": torch.tensor(float(self._step_dev.item()))}

    def state_dict(self):
        sd = super().state_dict()
        if self._seg is not None:       # per-parameter step counts, as torch.optim.AdamW keeps them
            idx = 0
            for gi, fg in enumerate(self._flat):
                steps = self._seg[gi][2].tolist()
                order = {id(p): i for i, p in enumerate(fg.params)}
                for p in self.param_groups[gi]["params"]:
                    j = order.get(id(p))
                    if j is not None and idx in sd["state"]:
                        sd["state"][idx]["step"] = torch.tensor(float(steps[j]))
                    idx += 1
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        if self._seg is not None:
            idx = 0
            for gi, fg in enumerate(self._flat):
                order = {id(p): i for i, p in enumerate(fg.params)}
                for p in self.param_groups[gi]["params"]:
                    j, entry = order.get(id(p)), state_dict["state"].get(idx)
                    if j is not None and entry is not None and "step" in entry:
                        self._seg[gi][2][j] = float(entry["step"])
                    idx += 1

    def _load_extra_state(self, extra):
        step = int(float(extra.get("step", 0)))
        self._steps = step
        self._step_dev.fill_(step)

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdamW.step does not take a closure")
        self._steps += 1
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
            self._check_grad_views()
        self._step_dev += 1          # a device op: replays of a captured step keep counting
        coef = None
        if self.max_grad_norm is not None:  # clip_grad_norm_ over ALL groups, computed and applied on device
            self._sumsq.zero_()
            for fg in self._flat:
                ops.sumsq(fg.flat_g, self._sumsq)
            ops.clip_coef(self._sumsq, self.max_grad_norm, self._coef, self.grad_norm)
            coef = self._coef
        want_lo = F.get_precision() == "fp32"
        for gi, (fg, g, m, v, hyper) in enumerate(zip(self._flat, self.param_groups, self._m, self._v, self._hyper_dev)):
            if want_lo:
                fg.ensure_lo()
            if self._seg is not None:
                seg_end, seg_flag, seg_step = self._seg[gi]
                ops.adamw_segments(fg.flat_p, fg.flat_g, m, v, g["eps"], hyper, seg_end, seg_flag, self._flags, seg_step,
                                   grad_scale=coef, shadow_hi=fg.flat_hi, shadow_lo=fg.flat_lo)
                fg.mark_fresh()
                continue
            ops.adamw(fg.flat_p, fg.flat_g, m, v, g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"],
                      self._steps, grad_scale=coef, shadow_hi=fg.flat_hi, shadow_lo=fg.flat_lo, hyper_dev=hyper,
                      step_dev=self._step_dev)
            fg.mark_fresh()
