"""L1 / L2 regularisers (reference: nn/regularizations.py:4-26).

`reg(weights) -> (loss, grad)` keeps the reference's call signature; the layer stack itself
uses `accumulate()` which adds the gradient into `param.grad` and the loss into a device
scalar in ONE kernel without any host read-back."""
from .._lib import REG_L1, REG_L2, lib
from .gpu import DeviceArray, LazyScalar, as_device, stream


class BaseRegularizer:
    kind = None

    def __init__(self, reg_strength):
        self.reg_strength = float(reg_strength)

    def accumulate(self, weights, grad, loss_dev):
        """grad += d reg/dw ; loss_dev[0] += reg(weights)   (BaseLayer.regularize, layers.py:147-155)"""
        lib.uocr_regularize(self.kind, weights.ptr, grad.ptr if grad is not None else None,
                            loss_dev.ptr if loss_dev is not None else None, weights.size,
                            self.reg_strength, stream())

    def __call__(self, weights):
        weights = as_device(weights)
        grad = DeviceArray.zeros(weights.shape)
        loss = DeviceArray.zeros((1,))
        self.accumulate(weights, grad, loss)
        return LazyScalar(loss), grad

    def __repr__(self):
        return f'{type(self).__name__}({self.reg_strength})'


class L1(BaseRegularizer):
    kind = REG_L1


class L2(BaseRegularizer):
    kind = REG_L2
