"""Optimisers (reference: nn/optimizers.py).  Per-parameter state keyed by id(param) as in the
reference (:14-29); `lr` is a public attribute read at every update (trainer.py:260 mutates it
per epoch).  Each update is ONE fused kernel over (w, g, state...) instead of the reference's
~9 elementwise CuPy kernels with temporaries (:56-61)."""
from .._lib import lib
from .gpu import DeviceArray, stream

EPS = 1e-8


class BaseOptimizer:
    state_names = ()

    def __init__(self):
        self.groups = {}

    def add_param(self, param):
        self.groups[id(param)] = (param, {})

    def _state(self, param):
        """Lazily allocated zero state tensors (the reference starts from scalar 0)."""
        if id(param) not in self.groups:
            self.add_param(param)
        state = self.groups[id(param)][1]
        for name, initial in zip(self.state_names, self.initials):
            cur = state.get(name)
            if cur is None or cur.shape != param.value.shape:
                state[name] = (DeviceArray.zeros(param.value.shape) if initial == 0
                               else DeviceArray.full(param.value.shape, initial))
        return state

    def update(self, param):
        raise NotImplementedError()


class Adam(BaseOptimizer):
    """v = b1 v + (1-b1) g; a = b2 a + (1-b2) g^2; w -= lr / (sqrt(a) + eps) * v.
    No bias correction, eps outside the sqrt -- optimizers.py:56-61."""
    state_names = ('velocity', 'accumulated')

    def __init__(self, lr=0.001, beta1=0.9, beta2=0.999, initial_velocity=0, initial_accumulated=0):
        super().__init__()
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2
        self.initials = [initial_velocity, initial_accumulated]

    def update(self, param, grad_scale=1.0, l2=0.0, reg_loss=None):
        st = self._state(param)
        lib.uocr_adam_update(param.value.ptr, param.grad.ptr, st['velocity'].ptr,
                             st['accumulated'].ptr, param.value.size, float(self.lr),
                             float(self.beta1), float(self.beta2), EPS, float(grad_scale), float(l2),
                             reg_loss.ptr if reg_loss is not None else None, stream())


class Momentum(BaseOptimizer):
    """optimizers.py:67-81"""
    state_names = ('velocity',)

    def __init__(self, lr, momentum=0, initial_velocity=0):
        super().__init__()
        self.lr, self.momentum = lr, momentum
        self.initials = [initial_velocity]

    def update(self, param):
        st = self._state(param)
        lib.uocr_momentum_update(param.value.ptr, param.grad.ptr, st['velocity'].ptr,
                                 param.value.size, float(self.lr), float(self.momentum), stream())


class RMSProp(BaseOptimizer):
    """optimizers.py:84-98"""
    state_names = ('accumulated',)

    def __init__(self, lr=0.01, rho=0.99, initial_accumulated=0):
        super().__init__()
        self.lr, self.rho = lr, rho
        self.initials = [initial_accumulated]

    def update(self, param):
        st = self._state(param)
        lib.uocr_rmsprop_update(param.value.ptr, param.grad.ptr, st['accumulated'].ptr,
                                param.value.size, float(self.lr), float(self.rho), EPS, stream())


class Adagrad(BaseOptimizer):
    """The reference's Adagrad reads `state.lr`, which never exists (optimizers.py:40), so it
    cannot run there; it is not used by my_model.  Kept only so the name resolves."""

    def __init__(self, lr=0.01, initial_accumulated=0):
        super().__init__()
        self.lr = lr
        self.initials = [initial_accumulated]

    def update(self, param):
        raise AttributeError("'State' object has no attribute 'lr'")   # reference behaviour
