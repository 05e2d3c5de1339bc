"""Host-side mirror of the reference's `web_app/components/nn` package for the B200 path.

Same class names, constructor arguments and method protocol as the reference (so a model built
for `nn` builds unchanged here), with every array living in device memory and every layer
forward/backward, loss, regulariser and optimiser step executed by hand-written sm_100a kernels
behind the C ABI of `include/uocr.h`.
"""
from . import gpu, help_func, initializers, layers, losses, models, optimizers, regularizations  # noqa: F401
from .gpu import CP, DeviceArray  # noqa: F401
