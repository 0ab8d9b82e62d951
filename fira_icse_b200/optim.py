"""Adam over flat buffers (the `optimizer.step()` of run_model.py:101-109, torch.optim.Adam semantics).

`FlatAdam(params, lr, groups=...)` re-homes the parameters it is given in ONE contiguous fp32 buffer (every
`p.data` becomes a view of it -- names, shapes, `state_dict()` and `load_state_dict()` are unchanged) and keeps
buffers of the same layout next to it: the gradients `g`, Adam's `m` and `v`, and a bf16 mirror of the parameters.
One `fira_adam_flat` launch updates everything (csrc/optim.cu) and refreshes the mirror.  What the layout buys:

  * the step is ONE memory-bound kernel instead of torch's multi-tensor launches over 264 tensors;
  * the bf16 GEMM operands of the throughput mode are views of the mirror: no per-step cast kernels;
  * `groups` of parameters laid out back to back (q|k, q|k|v, the 12 cross-attention k|v projections) make the
    concatenated weights of the fused projections plain views (`cat_rows`): no per-step `torch.cat`;
  * backward passes write parameter gradients straight into views of `g` (`grad_dest`), so neither the optimizer nor
    the data-parallel all-reduce needs a packing pass: the all-reduce runs on `g` itself.
"""
import weakref

import torch

from . import _lib

ALIGN = 64          # elements: every tensor starts on a 256-byte (fp32) / 128-byte (bf16) boundary


class _Slot:
    __slots__ = ("owner", "off", "numel")

    def __init__(self, owner, off, numel):
        self.owner, self.off, self.numel = weakref.ref(owner), off, numel


_SLOTS = {}          # (data_ptr, numel) of a re-homed parameter or of a cat_rows view -> _Slot


def _lookup(t):
    key = (t.data_ptr(), t.numel())
    s = _SLOTS.get(key)
    if s is None:
        return None, None
    o = s.owner()
    if o is None or o.p.data_ptr() + 4 * s.off != key[0]:
        del _SLOTS[key]                                       # the optimizer is gone: the address means nothing any more
        return None, None
    return o, s


def mirror_of(t):
    """bf16 view of a re-homed parameter / cat_rows view while the mirror is fresh, else None"""
    o, s = _lookup(t)
    if o is None or not o.fresh:
        return None
    return o.p16[s.off:s.off + s.numel].view(t.shape)


def grad_dest(ts, shape):
    """A NEW view, of shape `shape`, of the flat gradient buffer over the parameters `ts` (which must lie back to back
    in one FlatAdam layout), or None.  Backward passes write gradients there and hand (slices of) the view to autograd,
    which adopts them as `.grad` -- no copy, no packing pass.  The buffer was zero-filled by zero_grad().  Each span is
    handed out once per zero_grad(): a second backward before the next zero_grad() gets None, allocates as usual and lets
    autograd accumulate."""
    o0, s0 = _lookup(ts[0])
    if o0 is None or not o0.direct:
        return None
    off = s0.off
    for t in ts:
        o, s = _lookup(t)
        if o is not o0 or s.off != off:
            return None
        off += s.numel
    if s0.off in o0.handed:
        return None
    o0.handed.add(s0.off)
    return o0.g[s0.off:off].view(shape)


def cat_rows(ts):
    """torch.cat(ts, 0) -- as a zero-copy view when the tensors are adjacent parameters of one flat buffer"""
    o0, s0 = _lookup(ts[0])
    if o0 is not None:
        off = s0.off
        for t in ts:
            o, s = _lookup(t)
            if o is not o0 or s.off != off or t.shape[1:] != ts[0].shape[1:]:
                break
            off += s.numel
        else:
            n = off - s0.off
            view = o0.p[s0.off:off].view((sum(t.shape[0] for t in ts),) + tuple(ts[0].shape[1:]))
            _SLOTS[(view.data_ptr(), n)] = _Slot(o0, s0.off, n)
            return view
    return torch.cat(ts, 0)


def plan_layout(params, groups=()):
    """Offsets (in elements) of `params` inside one flat buffer and the buffer length: the members of every group whose
    parameters are all present lie back to back, in the group's order; every group / ungrouped tensor starts on an
    ALIGN-element boundary; the length is a multiple of ALIGN.  Pure host logic (works on any tensors)."""
    mine = {id(p) for p in params}
    order, seen = [], set()
    for grp in groups:                                   # adjacency groups first, members back to back
        if all(id(p) in mine for p in grp) and not any(id(p) in seen for p in grp):
            order.append(list(grp))
            seen.update(id(p) for p in grp)
    for p in params:
        if id(p) not in seen:
            order.append([p])
            seen.add(id(p))
    offs, off = {}, 0
    for grp in order:
        off = (off + ALIGN - 1) // ALIGN * ALIGN
        for p in grp:
            if len(grp) > 1 and p.numel() % 8:
                raise ValueError("FlatAdam: grouped parameters must have a multiple of 8 elements")
            offs[id(p)] = off
            off += p.numel()
    return [offs[id(p)] for p in params], (off + ALIGN - 1) // ALIGN * ALIGN


class FlatAdam:
    """torch.optim.Adam(params, lr, betas, eps) on flat buffers; `step()` is one kernel launch (capturable)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, groups=()):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.FiraLibraryError("FlatAdam: parameters must live on a CUDA device (no CPU path)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.offsets, self.n = plan_layout(self.params, groups)
        f32 = dict(dtype=torch.float32, device=dev)
        self.p = torch.zeros(self.n, **f32)
        self.g = torch.zeros(self.n, **f32)
        self.m = torch.zeros(self.n, **f32)
        self.v = torch.zeros(self.n, **f32)
        self.p16 = torch.zeros(self.n, dtype=torch.bfloat16, device=dev)
        self.step_t = torch.zeros((), **f32)
        self.grad_scale = None                                # device fp32 scalar: gradients are divided by it
        self.fresh = False                                    # the bf16 mirror equals the parameters
        self.direct = False                                   # g is zero-filled and may be written by backward passes
        self.handed = set()
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                _SLOTS[(view.data_ptr(), p.numel())] = _Slot(self, o, p.numel())
        self.gviews = [self.g[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]
        self.sync_mirror()

    def sync_mirror(self):
        """bf16 mirror <- parameters (after load_state_dict / any update that did not go through step())"""
        _lib.call("fira_cast_bf16", self.p.data_ptr(), self.p16.data_ptr(), self.n, torch.cuda.current_stream().cuda_stream)
        self.fresh = True

    # ------------------------------------------------------------------ torch.optim surface
    def zero_grad(self, set_to_none=True):
        """`.grad = None` for every parameter and ONE zero-fill of the flat gradient buffer: backward passes then write
        (or atomically accumulate) their results into it directly (grad_dest)"""
        for p in self.params:
            p.grad = None
        self.g.zero_()
        self.direct = True
        self.handed.clear()

    def gather_grads(self):
        """Gradients the backward did not write into `g` itself (ops.grad_dest) are copied there; returns how many."""
        src, dst = [], []
        base = self.p.data_ptr()
        for p, gv, o in zip(self.params, self.gviews, self.offsets):
            if p.data_ptr() != base + 4 * o:
                raise RuntimeError("FlatAdam: a parameter no longer lives in the flat buffer (model.to()/deepcopy after the "
                                   "optimizer was built?) -- build the optimizer after moving the model")
            if p.grad is None:
                raise RuntimeError("FlatAdam.step: a parameter received no gradient (torch.optim.Adam would skip it; "
                                   "leave dead parameters out of the optimizer instead)")
            if p.grad.data_ptr() != gv.data_ptr():
                src.append(p.grad.reshape(gv.shape))
                dst.append(gv)
        if src:
            torch._foreach_copy_(dst, src)
        return len(src)

    def step(self):
        self.gather_grads()
        self.direct = False                                   # g now holds this step's gradients
        self.step_t.add_(1.0)
        _lib.call("fira_adam_flat", self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                  self.p16.data_ptr(), self.n, self.lr, self.betas[0], self.betas[1], self.eps, self.step_t.data_ptr(),
                  self.grad_scale.data_ptr() if self.grad_scale is not None else None,
                  torch.cuda.current_stream().cuda_stream)
        self.fresh = True

    def _state_views(self, buf):
        return [buf[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def state_dict(self):
        return {"step": float(self.step_t.item()), "lr": self.lr, "betas": self.betas, "eps": self.eps,
                "exp_avg": [t.clone() for t in self._state_views(self.m)],
                "exp_avg_sq": [t.clone() for t in self._state_views(self.v)]}

    def load_state_dict(self, sd):
        self.step_t.fill_(float(sd["step"]))
        for dst, src in zip(self._state_views(self.m), sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self._state_views(self.v), sd["exp_avg_sq"]):
            dst.copy_(src)


def attach(model, optims):
    """Tell `model` which FlatAdam instances own its parameters: TransModel.forward refreshes a stale bf16 mirror
    before it runs, and load_state_dict marks the mirrors stale."""
    model._flat_optims = list(optims)
    if not getattr(model, "_flat_hooked", False):
        model._flat_hooked = True

        def _stale(module, incompatible_keys):
            for o in getattr(module, "_flat_optims", ()):
                o.fresh = False
        model.register_load_state_dict_post_hook(_stale)


def ensure_fresh(model):
    for o in getattr(model, "_flat_optims", ()):
        if not o.fresh:
            o.sync_mirror()
