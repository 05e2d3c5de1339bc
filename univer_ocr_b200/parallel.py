"""Batch-sharded data-parallel training step and the fused parameter update.

The reference trains with batch 1 on one device and, per step, runs for EVERY parameter tensor
`regularizer(w)` (+ a host sync for the loss), `grad += reg_grad`, then ~9 CuPy kernels of Adam
(`nn/layers/layers.py:147-155`, `nn/optimizers.py:56-61`), then re-allocates zero gradients
(`layers.py:20-21`).  Here all parameters of a model live in ONE flat device buffer (values,
gradients, Adam velocity / accumulator), every `Param.value / .grad` is a view into it, and one
step is

    forward + loss + backward (gradients accumulate into the flat buffer)
    [world > 1]  ONE NCCL sum-allreduce over the flat gradient buffer (NVLink / NVSwitch)
    ONE fused kernel per regularisation group: g*scale + 2*l2*w -> Adam -> w   (+ reg loss)
    ONE memset of the flat gradient buffer

Gradient scaling: the Dice / Jaccard losses sum over the batch (`losses.py:23`), so shard
gradients are summed without scaling; SoftmaxCE / SigmoidCE divide by the LOCAL batch
(`losses.py:69-72`), so their summed gradients are scaled by 1 / world (= local / global batch).
The L2 gradient depends on the weights only, so it is added once, after the allreduce.
"""
import numpy as np

from .nn import optimizers
from .nn.gpu import CP, DeviceArray, LazyScalar, stream
from .nn.losses import SegmentationDice2D, SegmentationJaccard2D
from .nn.regularizations import L2
from ._lib import lib


class FlatParameters:
    """Moves a model's parameters into flat buffers, grouped by L2 strength so that each group
    is one contiguous range (one fused update launch)."""

    def __init__(self, model):
        self.model = model
        groups = {}
        for lname, layer in model.layers.items():
            reg = getattr(layer, 'regularizer', None)
            if reg is not None and not isinstance(reg, L2):
                raise NotImplementedError('fused update supports L2 or no regulariser')
            l2 = float(reg.reg_strength) if reg is not None else 0.0
            if not layer.trainable:
                continue
            for pname, param in layer.params().items():
                groups.setdefault(l2, []).append((f'{lname}/{pname}', param))
        self.entries = []                       # (key, param, offset, size)
        self.groups = []                        # (l2, offset, size)
        offset = 0
        for l2 in sorted(groups, reverse=True):
            start = offset
            for key, param in groups[l2]:
                size = param.value.size
                size_al = (size + 3) // 4 * 4   # keep every tensor 16-byte aligned
                self.entries.append((key, param, offset, size))
                offset += size_al
            self.groups.append((l2, start, offset - start))
        self.total = offset
        self.values = DeviceArray.zeros((self.total,))
        self.grads = DeviceArray.zeros((self.total,))
        self.velocity = DeviceArray.zeros((self.total,))
        self.accumulated = DeviceArray.zeros((self.total,))
        self.adopt()

    def adopt(self):
        """(Re-)installs the views; call again after `set_weights` replaced parameter tensors."""
        for key, param, offset, size in self.entries:
            view = self.values.flat_view(offset, size, param.value.shape)
            if param.value.ptr != view.ptr:
                lib.uocr_memcpy_d2d(view.ptr, param.value.ptr, view.nbytes, stream())
                param._value = view
            gview = self.grads.flat_view(offset, size, param.value.shape)
            if param.grad.ptr != gview.ptr:
                param.grad = gview
        lib.uocr_memset(self.grads.ptr, 0, self.grads.nbytes, stream())


class DataParallel:
    """One training step of `model` on this rank's shard of the batch.

        dp = DataParallel(model, optimizer)        # world / rank from torch.distributed if initialised
        losses = dp.train(X_shard, y_shard)        # same dict as Model.train

    With world == 1 this is simply the fused single-GPU step."""

    def __init__(self, model, optimizer=None, process_group=None):
        self.model = model
        model.compute_input_grads = False            # a training step never reads dL/dX
        self.flat = FlatParameters(model)
        self.optimizer = optimizer if optimizer is not None else self._find_optimizer(model)
        if not isinstance(self.optimizer, optimizers.Adam):
            raise NotImplementedError('DataParallel fuses the Adam update (the only optimiser my_model uses)')
        self.world, self.rank, self._torch, self._dist, self._tensor = 1, 0, None, None, None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                import torch
                self._torch, self._dist = torch, dist
                self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
                self.group = process_group
        except ImportError:
            pass
        losses = model.loss if isinstance(model.loss, list) else [model.loss]
        summed = all(isinstance(l, (SegmentationDice2D, SegmentationJaccard2D)) for l in losses)
        self.grad_scale = 1.0 if summed else 1.0 / self.world
        if self.world > 1:
            self.broadcast_parameters()

    @staticmethod
    def _find_optimizer(model):
        for layer in model.layers.values():
            if layer.params():
                return layer.optimizer
        raise ValueError('model has no parameters')

    def _as_tensor(self, arr):
        """Zero-copy torch view of a DeviceArray (through __cuda_array_interface__)."""
        return self._torch.as_tensor(arr, device=f'cuda:{self._torch.cuda.current_device()}')

    def _on_compute_stream(self):
        return self._torch.cuda.stream(self._torch.cuda.ExternalStream(stream()))

    def broadcast_parameters(self, src=0):
        with self._on_compute_stream():
            self._dist.broadcast(self._as_tensor(self.flat.values), src=src, group=self.group)

    def allreduce_gradients(self):
        if self.world == 1:
            return
        if self._tensor is None:
            self._tensor = self._as_tensor(self.flat.grads)
        with self._on_compute_stream():        # NCCL orders itself after the backward kernels
            self._dist.all_reduce(self._tensor, op=self._dist.ReduceOp.SUM, group=self.group)

    def update(self):
        """Fused L2 + Adam over the flat buffers, then zero the gradients.  Returns the
        regularisation loss (of the pre-update weights, like `Model.regularize`)."""
        opt, flat = self.optimizer, self.flat
        reg_loss = DeviceArray.zeros((1,))
        for l2, offset, size in flat.groups:
            if size == 0:
                continue
            lib.uocr_adam_update(flat.values.ptr + 4 * offset, flat.grads.ptr + 4 * offset,
                                 flat.velocity.ptr + 4 * offset, flat.accumulated.ptr + 4 * offset, size,
                                 float(opt.lr), float(opt.beta1), float(opt.beta2), optimizers.EPS,
                                 float(self.grad_scale), l2, reg_loss.ptr if l2 else None, stream())
        lib.uocr_memset(flat.grads.ptr, 0, flat.grads.nbytes, stream())
        CP.weights_generation += 1                      # the parameter views changed under the layers
        return LazyScalar(reg_loss)

    def train(self, X, y):
        model = self.model
        X = X if isinstance(X, list) else [X]
        y = y if isinstance(y, list) else [y]
        predicted = model.forward(X, clear_grads=False)
        losses, gradients = [], []
        for key in range(model.outputs_count):
            loss, grad = model._loss_for(key)(predicted[key], y[key])
            losses.append(loss)
            gradients.append(grad)
        model.backward(gradients)
        self.allreduce_gradients()
        reg = self.update()
        return {'output_losses': losses, 'regularization_loss': reg}
