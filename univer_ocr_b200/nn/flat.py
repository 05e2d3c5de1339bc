"""Flat parameter storage and the fused parameter update.

The reference runs, per step and for EVERY parameter tensor, `regularizer(w)` (+ a host sync for
the loss), `grad += reg_grad`, then ~9 CuPy kernels of Adam (`nn/layers/layers.py:147-155`,
`nn/optimizers.py:56-61`), then re-allocates zero gradients (`layers.py:20-21`).  Here all trainable
parameters of a model live in ONE flat device buffer (values, gradients, Adam velocity /
accumulator); every `Param.value / .grad` and the optimiser's per-parameter state are views into it,
and the update of a step is

    ONE fused kernel per regularisation group:  g*scale + 2*l2*w -> Adam -> w   (+ reg loss)
    ONE memset of the flat gradient buffer

`Model.train` takes this route whenever the model qualifies (`FlatParameters.eligible`);
`parallel.DataParallel` adds the gradient allreduce between backward and update.  Because the views
are shared, the per-parameter protocol of the reference (`param.update_grad()`, `layer.regularize()`,
`model.set_weights(...)`) keeps working on the same memory.
"""
from .._lib import lib
from . import optimizers
from .gpu import CP, DeviceArray, LazyScalar, stream
from .regularizations import L2


class FlatParameters:
    """Moves a model's trainable parameters into flat buffers, grouped by L2 strength so that each
    group is one contiguous range (one fused update launch); within a group, layers keep the model's
    evaluation order and a layer's tensors are adjacent (`layer_ranges`: what a gradient bucket sends)."""

    @staticmethod
    def eligible(model):
        """The single shared Adam instance if `model` can use the fused update, else None: every layer that owns
        parameters is trainable, regularised by L2 or nothing, and updated by ONE plain `Adam` with zero initial
        state (the reference's own configuration, my_model/train.py:127, my_model/model.py:37-39)."""
        if not getattr(model, 'trainable', True):
            return None
        opt = None
        for layer in model.layers.values():
            params = layer.params()
            if not params:
                continue
            reg = getattr(layer, 'regularizer', None)
            if not layer.trainable or (reg is not None and type(reg) is not L2):
                return None
            for param in params.values():
                if opt is None:
                    opt = param.optimizer
                if param.optimizer is not opt:
                    return None
        if type(opt) is not optimizers.Adam or list(opt.initials) != [0, 0]:
            return None
        return opt

    def __init__(self, model):
        self.model = model
        groups, order = {}, list(getattr(model, '_order', None) or model.layers)
        for lname in order + [n for n in model.layers if n not in order]:
            layer = model.layers[lname]
            reg = getattr(layer, 'regularizer', None)
            if reg is not None and not isinstance(reg, L2):
                raise NotImplementedError('fused update supports L2 or no regulariser')
            l2 = float(reg.reg_strength) if reg is not None else 0.0
            if not layer.trainable:
                continue
            for pname, param in layer.params().items():
                groups.setdefault(l2, []).append((lname, f'{lname}/{pname}', param))
        self.entries = []                       # (key, param, offset, size)
        self.groups = []                        # (l2, offset, size)
        self.layer_ranges = {}                  # layer name -> (offset, end), 16-byte aligned
        offset = 0
        for l2 in sorted(groups, reverse=True):
            start = offset
            for lname, key, param in groups[l2]:
                size = param.value.size
                size_al = (size + 3) // 4 * 4   # keep every tensor 16-byte aligned
                self.entries.append((key, param, offset, size))
                lo, _ = self.layer_ranges.get(lname, (offset, offset))
                self.layer_ranges[lname] = (lo, offset + size_al)
                offset += size_al
            self.groups.append((l2, start, offset - start))
        self.total = offset
        self.values = DeviceArray.zeros((self.total,))
        self.grads = DeviceArray.zeros((self.total,))
        self.velocity = DeviceArray.zeros((self.total,))
        self.accumulated = DeviceArray.zeros((self.total,))
        self.clean = False                      # True while the flat gradient buffer is known to be all zero
        self.adopt()

    def adopt(self):
        """(Re-)installs the views.  Parameter tensors, gradients and Adam state that live elsewhere are copied in
        first, so adopting a model in the middle of a run loses nothing."""
        for key, param, offset, size in self.entries:
            shape = param.value.shape
            view = self.values.flat_view(offset, size, shape)
            if param.value.ptr != view.ptr:
                lib.uocr_memcpy_d2d(view.ptr, param.value.ptr, view.nbytes, stream())
            gview = self.grads.flat_view(offset, size, shape)
            param.pin(view, gview)
            opt = param.optimizer
            state = opt.groups[id(param)][1] if id(param) in getattr(opt, 'groups', {}) else None
            if isinstance(opt, optimizers.Adam):
                views = {'velocity': self.velocity.flat_view(offset, size, shape),
                         'accumulated': self.accumulated.flat_view(offset, size, shape)}
                for name, v in views.items():
                    old = state.get(name) if state else None
                    if old is not None and old.ptr != v.ptr and old.shape == shape:
                        lib.uocr_memcpy_d2d(v.ptr, old.ptr, v.nbytes, stream())
                opt.groups[id(param)] = (param, views)
        lib.uocr_memset(self.grads.ptr, 0, self.grads.nbytes, stream())
        self.clean = True
        CP.weights_generation += 1

    def attached(self):
        """False if some parameter no longer lives in the flat buffers (someone re-bound `Param._value`)."""
        return all(param.value.ptr == self.values.ptr + 4 * offset for _, param, offset, _ in self.entries)

    def zero_grads(self):
        lib.uocr_memset(self.grads.ptr, 0, self.grads.nbytes, stream())
        self.clean = True

    def update(self, opt, grad_scale=1.0):
        """Fused L2 + Adam over the flat buffers, then zero the gradients.  Returns the regularisation loss (of
        the pre-update weights, like `Model.regularize`) as a LazyScalar, or 0 if nothing is regularised."""
        if not self.attached():
            self.adopt()
        regularised = any(l2 for l2, _, size in self.groups if size)
        reg_loss = DeviceArray.zeros((1,)) if regularised else None
        for l2, offset, size in self.groups:
            if size == 0:
                continue
            lib.uocr_adam_update(self.values.ptr + 4 * offset, self.grads.ptr + 4 * offset,
                                 self.velocity.ptr + 4 * offset, self.accumulated.ptr + 4 * offset, size,
                                 float(opt.lr), float(opt.beta1), float(opt.beta2), optimizers.EPS,
                                 float(grad_scale), l2, reg_loss.ptr if l2 else None, stream())
        self.zero_grads()
        CP.weights_generation += 1                      # the parameter views changed under the layers
        return LazyScalar(reg_loss) if regularised else 0

    # ---- roll-back snapshots (Trainer) ----------------------------------------------
    def snapshot(self):
        return self.values.copy()

    def restore(self, snap):
        lib.uocr_memcpy_d2d(self.values.ptr, snap.ptr, self.values.nbytes, stream())
        CP.weights_generation += 1


class BucketScheduler:
    """Which contiguous ranges of the flat gradient buffer to allreduce, and when, while backward runs.

    Backward finishes layers in reverse evaluation order; within a regularisation group the flat layout follows
    the evaluation order, so finished ranges grow downwards and stay contiguous.  A bucket is flushed as soon as
    it holds `bucket_elems` gradients, when the next finished layer is not adjacent (group switch), or at the
    end of backward -- so the Char head's FullyConnected gradients (99 % of my_model's 3.2 MB) are on the wire
    while the convolutions' backward is still running.  Pure bookkeeping (unit-tested on CPU)."""

    def __init__(self, layer_ranges, bucket_elems=1 << 16):
        self.layer_ranges = dict(layer_ranges)
        self.bucket_elems = int(bucket_elems)
        self.pending = None
        self.done = set()

    def reset(self):
        self.pending, self.done = None, set()

    def layer_done(self, name):
        """-> list of (lo, hi) ranges to send now that `name`'s parameter gradients are final."""
        rng = self.layer_ranges.get(name)
        if rng is None or name in self.done or rng[0] == rng[1]:
            return []
        self.done.add(name)
        out = []
        lo, hi = rng
        if self.pending is None:
            self.pending = (lo, hi)
        elif hi == self.pending[0]:
            self.pending = (lo, self.pending[1])
        elif lo == self.pending[1]:
            self.pending = (self.pending[0], hi)
        else:
            out.append(self.pending)
            self.pending = (lo, hi)
        if self.pending[1] - self.pending[0] >= self.bucket_elems:
            out.append(self.pending)
            self.pending = None
        return out

    def finish(self):
        """-> what is left at the end of backward, including layers that never reported (defensive)."""
        out = []
        if self.pending is not None:
            out.append(self.pending)
            self.pending = None
        for name, (lo, hi) in self.layer_ranges.items():
            if name not in self.done and hi > lo:
                out.append((lo, hi))
                self.done.add(name)
        return out
