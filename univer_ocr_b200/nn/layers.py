"""Layer stack of the B200 path -- same classes, constructor arguments and method protocol as
the reference's `nn/layers/{layers,convolutional,maxpool,upsample}.py`, with the arithmetic in
libuocr (include/uocr.h).  Protocol kept from `BaseLayer` (layers.py:24-166):

    initialize(input_shapes)            forward(list-or-array) -> list
    backward(list-or-array) -> list     params() -> {'w': Param, ...}
    get_weights() / set_weights(dict)   clear_grads() / update_grads() / regularize()
    nan_weights() / count_parameters()  get_output_shapes() / get_all_output_shapes()
    _get_receptive_field() / is_fully_convolutional() / changes_receptive_field()
    init_progress_tracker() / _set_name()

Differences that are deliberate: arrays are float32 DeviceArrays; the saved
activations are the *unpadded* inputs (padding is synthesised inside the kernels);
`regularize()` returns a LazyScalar; nothing synchronises the device per layer.
"""
import ctypes
import math

import numpy as np

from .._lib import ACT_LEAKY, ACT_NONE, ACT_SIGMOID, MATH_TF32, ConvDesc, lib
from .gpu import CP, DeviceArray, LazyScalar, as_device, concatenate, slice_axis, stream
from .help_func import make_list_if_not, tuplize
from .initializers import kaiming_uniform
from .optimizers import Adam
from .progress_tracker import BaseProgressTracker, track_method

class UpsampledInput:
    """Saved-for-backward input of a Convolutional2D that ran with a folded Upsample2D(2) in front:
    the LOW-resolution tensor (what is kept in memory) and the upsampling factor."""

    def __init__(self, array, factor):
        self.array, self.factor = array, factor


class FromOutput:
    """Marker stored in an activation layer's memory when the activation ran as the epilogue
    of the producing kernel: only its OUTPUT exists, and the backward is evaluated from it."""
    __slots__ = ('y',)

    def __init__(self, y):
        self.y = y


def _kmajor_copy(layer, w, k_rows, n_cols):
    """K-major (n_cols, k_rows) copy of the layer's weight matrix `w` (k_rows [+ bias row], n_cols) for the
    tensor-core kernels, cached on the layer until any parameter changes (CP.weights_generation)."""
    cached = getattr(layer, '_kmajor_cache', None)
    if cached is not None and cached[0] == CP.weights_generation and cached[1] == w.ptr:
        return cached[2]
    wt = DeviceArray((n_cols, k_rows))
    lib.uocr_weights_to_kmajor(w.ptr, wt.ptr, k_rows, n_cols, stream())
    layer._kmajor_cache = (CP.weights_generation, w.ptr, wt)
    return wt


_DEFAULT_OPTIMIZER = Adam()     # the reference's default argument is one shared instance (layers.py:31)


class Param:
    """A trainable tensor with its gradient and optimiser hook (layers.py:10-21).

    Once `nn.flat.FlatParameters` has adopted the parameter (`pin`), `value` and `grad` are views into the model's
    flat buffers and assignments copy INTO those views instead of re-binding the attribute, so `set_weights`,
    roll-backs and user code that writes `param.grad = ...` keep operating on the memory the fused update reads."""

    def __init__(self, value, optimizer=None):
        self._value = None
        self._grad = None
        self._pinned = False
        self._pending = None                 # host values not uploaded yet (see `value`)
        self.value = value
        self.optimizer = optimizer
        self.optimizer.add_param(self)

    @staticmethod
    def _coerce(new):
        return new if isinstance(new, DeviceArray) else CP.copy(np.asarray(new, dtype=np.float64))

    def _assign(self, slot, new):
        cur = getattr(self, slot)
        new = self._coerce(new)
        if self._pinned and cur is not None and new.shape == cur.shape:
            if new.ptr != cur.ptr:
                lib.uocr_memcpy_d2d(cur.ptr, new.ptr, cur.nbytes, stream())
        else:
            if self._pinned and slot == '_value':
                self._pinned = False                    # shape changed: the flat owner re-adopts on its next update
            setattr(self, slot, new)

    # Host values handed to a parameter that has no device tensor yet (construction: initialiser draws, `w=` / `b=`
    # arguments) are uploaded on first use, and the zero gradient is allocated on first use: building a network and
    # analysing it (`get_all_output_shapes`, `count_parameters`, receptive fields, fusion plans) needs no device.
    @property
    def shape(self):
        return self._pending.shape if self._pending is not None else self._value.shape

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def value(self):
        if self._pending is not None:
            host, self._pending = self._pending, None
            self._value = CP.copy(host)
        return self._value

    @value.setter
    def value(self, new):
        if self._value is None and not isinstance(new, DeviceArray):
            self._pending = np.asarray(new, dtype=np.float64)
        else:
            self._pending = None
            self._assign('_value', new)
        CP.weights_generation += 1

    @property
    def grad(self):
        if self._grad is None:
            self._grad = DeviceArray.zeros(self.shape)
        return self._grad

    @grad.setter
    def grad(self, new):
        self._assign('_grad', new)

    def pin(self, value_view, grad_view):
        """Re-binds value / grad to views of a flat buffer (their contents were copied by the caller)."""
        self._value, self._grad, self._pinned = value_view, grad_view, True

    def update_grad(self):
        self.optimizer.update(self)
        CP.weights_generation += 1

    def clear_grad(self):
        if self._grad is None:
            return                               # allocated as zeros on first use
        if self._grad.shape != self.shape:
            self._grad = DeviceArray.zeros(self.shape)
        else:
            self._grad.fill(0)


class BaseLayer:
    def __init__(self, name=None, input_shapes=None, trainable=True, initializer=kaiming_uniform,
                 regularizer=None, optimizer=_DEFAULT_OPTIMIZER):
        self.name = name
        self.input_shapes = input_shapes
        self.inputs_count = len(input_shapes) if input_shapes is not None else None
        self.trainable = trainable
        self.initializer = initializer
        self.regularizer = regularizer
        self.optimizer = optimizer
        self.is_initialized = True
        self._mem = {}
        self._receptive_fields = {}
        self.progress_tracker = BaseProgressTracker()

    # ---- lifecycle ----------------------------------------------------------------
    def initialize_from_X(self, X):
        self.initialize([x.shape for x in make_list_if_not(X)])

    def initialize(self, input_shapes):
        self.input_shapes = input_shapes
        self.inputs_count = len(input_shapes)
        self.is_initialized = True

    @track_method('forward')
    def forward(self, inputs):
        assert self.is_initialized, 'You must initialize() layer before calling forward() method'
        return [self._forward(as_device(X), mem_id)
                for mem_id, X in enumerate(make_list_if_not(inputs))]

    @track_method('backward')
    def backward(self, grads):
        result = [self._backward(as_device(grad), mem_id)
                  for mem_id, grad in enumerate(make_list_if_not(grads))]
        self.clear_memory()
        return result

    def _forward(self, X, mem_id=0):
        raise NotImplementedError()

    def _backward(self, grad, mem_id=0):
        raise NotImplementedError()

    def clear_memory(self):
        self._mem = {}

    # ---- parameters ---------------------------------------------------------------
    def params(self):
        return {}

    def update_grads(self):
        if not self.trainable:
            return
        for param in self.params().values():
            param.update_grad()

    def clear_grads(self):
        for param in self.params().values():
            param.clear_grad()

    def get_weights(self):
        return {name: param.value.tolist() for name, param in self.params().items()}

    def set_weights(self, weights):
        """Skips missing keys, NaN tensors and shape mismatches with the reference's messages
        (layers.py:123-137)."""
        for name, param in self.params().items():
            new = weights.get(name, None)
            if new is None:
                continue
            new = np.array(new)
            error = None
            if np.any(np.isnan(new)):
                error = 'NaN found in loaded weights'
            elif new.shape != param.shape:
                error = f'Shapes don`t match: {new.shape} != {param.shape}'
            if error is not None:
                print(f'{self.name}/{name}: {error}, skipping')
                continue
            param.value = new                    # uploaded now if the parameter already lives on the device

    def nan_weights(self):
        return any(param.value.isnan_any() for param in self.params().values())

    def count_parameters(self, param=None):
        if param is not None:
            return self.params()[param].size
        return sum(p.size for p in self.params().values())

    def regularize(self, loss_dev=None):
        """grad += regulariser gradient for EVERY param of the layer (w and b), returns the
        regulariser loss (layers.py:147-155).  With `loss_dev` the loss is accumulated into that
        device scalar (used by Model.regularize) and nothing is returned to the host."""
        if self.regularizer is None:
            return 0
        own = loss_dev is None
        if own:
            loss_dev = DeviceArray.zeros((1,))
        for param in self.params().values():
            self.regularizer.accumulate(param.value, param.grad, loss_dev)
        return LazyScalar(loss_dev) if own else 0

    # ---- shapes / receptive fields ------------------------------------------------
    def get_all_output_shapes(self, input_shapes):
        return self.get_output_shapes(input_shapes), {}

    def get_output_shapes(self, input_shapes):
        raise NotImplementedError()

    def get_outputs_count(self):
        return 1

    def is_fully_convolutional(self):
        return True

    def changes_receptive_field(self):
        return False

    def _get_receptive_field(self, axis, position, output_id):
        assert output_id < self.get_outputs_count(), (
            f'This layer has only {self.get_outputs_count()} outputs')
        return {0: {position}}

    def _window_receptive_field(self, axis, position, output_id, kind):
        """Shared by Convolutional2D / MaxPool2D: taps position*s - p + [0, k)."""
        assert 0 <= axis < 2, f'{kind} has two axis, found {axis}'
        assert output_id < self.get_outputs_count(), (
            f'This layer has only {self.get_outputs_count()} outputs')
        key = (axis, position, output_id)
        if key not in self._receptive_fields:
            start = position * self.stride[axis] - self.padding[axis]
            self._receptive_fields[key] = {0: set(range(start, start + self.kernel_size[axis]))}
        return self._receptive_fields[key]

    def _clear_receptive_fields_info(self):
        self._receptive_fields = {}

    # ---- misc ---------------------------------------------------------------------
    def _set_name(self, name):
        self.name = name

    def _init_optimizer(self):
        for param in self.params().values():
            self.optimizer.add_param(param)

    def init_progress_tracker(self, progress_tracker, set_names_recursively=False):
        self.progress_tracker = progress_tracker
        self.progress_tracker.register_layer(self.name)


BaseLayerGPU = BaseLayer     # the reference's CPU/GPU dispatch subclass (layers.py:169-237) collapses


# ------------------------------------------------------------------------------------------
# structural layers
# ------------------------------------------------------------------------------------------

class Concat(BaseLayer):
    """layers.py:240-284"""

    def __init__(self, axis=-1, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.axis = axis
        self.is_initialized = self.inputs_count is not None

    @track_method('forward')
    def forward(self, inputs):
        if not isinstance(inputs, list):
            self._mem = [inputs.shape]
            return inputs
        inputs = [as_device(X) for X in inputs]
        self._mem = [X.shape for X in inputs]
        return [concatenate(inputs, self.axis)]

    @track_method('backward')
    def backward(self, grads):
        grad = as_device(make_list_if_not(grads)[0])
        result, pos = [], 0
        for shape in self._mem:
            width = shape[self.axis]
            result.append(slice_axis(grad, self.axis, pos, pos + width))
            pos += width
        self.clear_memory()
        return result

    def get_output_shapes(self, input_shapes):
        input_shapes = make_list_if_not(input_shapes)
        out = list(input_shapes[0])
        axis = self.axis % len(out)
        assert axis != 0 or len(input_shapes) == 1
        out[axis] = sum(shape[axis] for shape in input_shapes)
        return [tuple(out)]

    def changes_receptive_field(self):
        return True

    def _get_receptive_field(self, axis, position, output_id):
        assert output_id < self.get_outputs_count(), (
            f'This layer has only {self.get_outputs_count()} outputs')
        return {in_key: {position} for in_key in range(self.inputs_count)}


class Flatten(BaseLayer):
    """layers.py:287-304"""

    def _forward(self, X, mem_id=0):
        self._mem[mem_id] = X.shape
        return X.reshape(self.get_output_shapes(X.shape)[0])

    def _backward(self, grad, mem_id=0):
        return grad.reshape(self._mem[mem_id])

    def get_output_shapes(self, input_shapes):
        shape = make_list_if_not(input_shapes)[0]
        return [(shape[0], int(np.prod(shape[1:])))]

    def is_fully_convolutional(self):
        return False

    def _get_receptive_field(self, axis, position, output_id):
        raise NotImplementedError('The method is not supported by Flatten Layer')


class Noop(BaseLayer):
    """layers.py:366-374"""

    def _forward(self, X, mem_id=0):
        return X

    def _backward(self, grad, mem_id=0):
        return grad

    def get_output_shapes(self, input_shapes):
        return make_list_if_not(input_shapes)


# ------------------------------------------------------------------------------------------
# activations
# ------------------------------------------------------------------------------------------

class LeakyRelu(BaseLayer):
    """y = X * ((X >= 0) + alpha * (X < 0)) (layers.py:390-404).  The mask is recomputed from
    the saved input in the backward kernel instead of being stored as a float tensor."""

    def __init__(self, alpha=0.01, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha = alpha

    def _forward(self, X, mem_id=0):
        self._mem[mem_id] = X
        y = DeviceArray(X.shape)
        lib.uocr_leaky_relu_fwd(X.ptr, y.ptr, X.size, float(self.alpha), stream())
        return y

    def _backward(self, grad, mem_id=0):
        X = self._mem[mem_id]
        if isinstance(X, FromOutput):
            dx = DeviceArray(X.y.shape)
            lib.uocr_act_bwd_from_output(X.y.ptr, grad.ptr, dx.ptr, dx.size, ACT_LEAKY,
                                         float(self.alpha), stream())
            return dx
        dx = DeviceArray(X.shape)
        lib.uocr_leaky_relu_bwd(X.ptr, grad.ptr, dx.ptr, X.size, float(self.alpha), stream())
        return dx

    def get_output_shapes(self, input_shapes):
        return make_list_if_not(input_shapes)


class Relu(LeakyRelu):
    """layers.py:377-387 (mask = X >= 0)."""

    def __init__(self, *args, **kwargs):
        super().__init__(0.0, *args, **kwargs)


class Sigmoid(BaseLayer):
    """layers.py:407-418"""

    def _forward(self, X, mem_id=0):
        self._mem[mem_id] = X
        y = DeviceArray(X.shape)
        lib.uocr_sigmoid_fwd(X.ptr, y.ptr, X.size, stream())
        return y

    @track_method('backward')
    def backward(self, grads):
        """A Dice / Jaccard gradient that has not been written out yet (losses.LazySegGrad) is combined with this
        layer's derivative in one pass; anything else takes the reference's route."""
        grads = make_list_if_not(grads)
        through = getattr(grads[0], 'through_sigmoid', None) if len(grads) == 1 else None
        X = self._mem.get(0)
        if through is not None and isinstance(X, DeviceArray) and X.shape == tuple(grads[0].shape) and len(X.shape) == 4:
            dx = through(X)
            if dx is not None:
                self.clear_memory()
                return [dx]
        result = [self._backward(as_device(grad), mem_id) for mem_id, grad in enumerate(grads)]
        self.clear_memory()
        return result

    def _backward(self, grad, mem_id=0):
        X = self._mem[mem_id]
        if isinstance(X, FromOutput):
            dx = DeviceArray(X.y.shape)
            lib.uocr_act_bwd_from_output(X.y.ptr, grad.ptr, dx.ptr, dx.size, ACT_SIGMOID, 0.0, stream())
            return dx
        dx = DeviceArray(X.shape)
        lib.uocr_sigmoid_bwd(X.ptr, grad.ptr, dx.ptr, X.size, stream())
        return dx

    def get_output_shapes(self, input_shapes):
        return make_list_if_not(input_shapes)


# ------------------------------------------------------------------------------------------
# FullyConnected
# ------------------------------------------------------------------------------------------

class FullyConnected(BaseLayer):
    """y = [X, 1] . W with the bias folded in as the last row of W (layers.py:307-363)."""

    def __init__(self, n_input=None, n_output=None, w=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_input, self.n_output, self.w = n_input, n_output, w
        if self.input_shapes is None and n_input is not None:
            self.input_shapes = [(None, self.n_input)]
        if self.input_shapes is not None:
            self.initialize(self.input_shapes)
        else:
            self.is_initialized = False

    def initialize(self, input_shapes):
        self.input_shapes = input_shapes
        self.n_input = self.input_shapes[0][1]
        if self.n_output is None:
            self.n_output = self.n_input
        if self.w is None:
            self.w = Param(self.initializer(self.n_input + 1, self.n_output), optimizer=self.optimizer)
        else:
            assert self.w.shape == (self.n_input + 1, self.n_output)
            self.w = Param(self.w if not isinstance(self.w, DeviceArray) else self.w.copy(),
                           optimizer=self.optimizer)
        self._init_optimizer()
        self.is_initialized = True

    def _forward(self, X, mem_id=0, act=ACT_NONE, alpha=0.0, in_upsample=1, save=True):
        """`act` / `alpha`: activation applied in the GEMM epilogue (Model fuses a following
        LeakyRelu / Sigmoid layer into this call)."""
        assert X.ndim == 2 and X.shape[1] == self.n_input, f'{X.shape} vs n_input={self.n_input}'
        assert in_upsample == 1
        if save:
            self._mem[mem_id] = X
        y = DeviceArray((X.shape[0], self.n_output))
        if not save and CP.math_mode == MATH_TF32:       # inference: weights are static, cache their K-major copy
            wt = _kmajor_copy(self, self.w.value, self.n_input, self.n_output)
            lib.uocr_fc_fwd_kmajor(X.ptr, self.w.value.ptr, wt.ptr, y.ptr, X.shape[0], self.n_input, self.n_output,
                                   act, float(alpha), CP.math_mode, stream())
        else:
            lib.uocr_fc_fwd(X.ptr, self.w.value.ptr, y.ptr, X.shape[0], self.n_input, self.n_output,
                            act, float(alpha), CP.math_mode, stream())
        return y

    def _backward(self, grad, mem_id=0):
        X = self._mem[mem_id]
        dx = DeviceArray(X.shape)
        lib.uocr_fc_bwd(X.ptr, self.w.value.ptr, grad.ptr, dx.ptr, self.w.grad.ptr, X.shape[0],
                        self.n_input, self.n_output, 1, CP.math_mode, stream())
        return dx

    def get_output_shapes(self, input_shapes):
        return [(make_list_if_not(input_shapes)[0][0], self.n_output)]

    def is_fully_convolutional(self):
        return False

    def changes_receptive_field(self):
        return True

    def _get_receptive_field(self, axis, position, output_id):
        raise NotImplementedError('The method is not supported by Fully Connected Layer')

    def params(self):
        return {'w': self.w}


# ------------------------------------------------------------------------------------------
# Convolutional2D + Conv2DToBatchedFixedWidthed
# ------------------------------------------------------------------------------------------

class Convolutional2D(BaseLayer):
    """NHWC convolution with padding / padding_value / stride / bias
    (layers/convolutional.py:12-327).  w: (kh, kw, Cin, Cout), b: (Cout,)."""

    def __init__(self, kernel_size, in_channels=None, out_channels=None, padding=0, padding_value=0,
                 stride=1, w=None, b=None, bias=True, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.kernel_size = tuplize('kernel_size', kernel_size, 2)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.padding = tuplize('padding', padding, 2)
        self.padding_value = padding_value
        self.stride = tuplize('stride', stride, 2)
        self.w, self.b, self.bias = w, b, bias
        if self.input_shapes is None and in_channels is not None:
            self.input_shapes = [(None, None, None, self.in_channels)]
        if self.input_shapes is not None:
            self.initialize(self.input_shapes)
        else:
            self.is_initialized = False

    def initialize(self, input_shapes):
        self.input_shapes = input_shapes
        self.in_channels = self.input_shapes[0][3]
        if self.out_channels is None:
            self.out_channels = self.in_channels
        w_shape = (*self.kernel_size, self.in_channels, self.out_channels)
        b_shape = (self.out_channels,)
        # one draw for [w; b] exactly like the reference (convolutional.py:41-45)
        wb = self.initializer(int(np.prod(w_shape[:3])) + 1, self.out_channels)
        w = np.reshape(wb[:-1, :], w_shape) if self.w is None else self.w
        b = np.reshape(wb[-1, :], b_shape) if self.b is None else self.b
        assert tuple(w.shape) == w_shape, f'{tuple(w.shape)} != {w_shape}'
        assert tuple(b.shape) == b_shape, f'{tuple(b.shape)} != {b_shape}'
        self.w = Param(w if not isinstance(w, DeviceArray) else w.copy(), optimizer=self.optimizer)
        self.b = Param(b if not isinstance(b, DeviceArray) else b.copy(), optimizer=self.optimizer)
        self._init_optimizer()
        self.is_initialized = True

    def _desc(self, x_shape, in_upsample=1):
        n, h, w, c = x_shape
        assert c == self.in_channels, f'input has {c} channels, layer expects {self.in_channels}'
        return ConvDesc(n, h * in_upsample, w * in_upsample, c, self.out_channels, *self.kernel_size,
                        *self.padding, *self.stride, float(self.padding_value), int(bool(self.bias)),
                        CP.math_mode, in_upsample)

    def _forward(self, X, mem_id=0, act=ACT_NONE, alpha=0.0, in_upsample=1, save=True):
        assert X.ndim == 4, f'expected NHWC input, got shape {X.shape}'
        desc = self._desc(X.shape, in_upsample)
        if save:
            assert in_upsample == 1 or self.supports_upsampled_input_grad()
            self._mem[mem_id] = X if in_upsample == 1 else UpsampledInput(X, in_upsample)
        n, h, w, c = X.shape
        y = DeviceArray(self.get_output_shapes((n, h * in_upsample, w * in_upsample, c))[0])
        if (not save and CP.math_mode == MATH_TF32 and in_upsample == 1 and self.in_channels % 32 == 0
                and self.out_channels % 16 == 0):        # a tcgen05 implicit-GEMM layer in inference
            kh, kw = self.kernel_size
            wt = _kmajor_copy(self, self.w.value, kh * kw * self.in_channels, self.out_channels)
            lib.uocr_conv2d_fwd_kmajor(ctypes.byref(desc), X.ptr, self.w.value.ptr, wt.ptr, self.b.value.ptr, y.ptr,
                                       act, float(alpha), stream())
        else:
            lib.uocr_conv2d_fwd(ctypes.byref(desc), X.ptr, self.w.value.ptr, self.b.value.ptr, y.ptr,
                                act, float(alpha), stream())
        return y

    def supports_upsampled_input_grad(self):
        """True if the weight gradient can read its input through a folded Upsample2D(2)
        : 5x5, stride 1, padding 2, 1 -> 1 channels (Paragraph) or 4 -> 4 / 2 channels (Line)."""
        if self.kernel_size != (5, 5) or self.stride != (1, 1) or self.padding != (2, 2):
            return False
        return ((self.in_channels == 1 and self.out_channels == 1)              # conv55_c1_wgrad_roll_kernel<UPS>
                or (self.in_channels == 4 and self.out_channels in (2, 4)       # conv55_c4_wgrad_tiled_kernel<UPS>
                    and self.padding_value == 0))

    def _backward(self, grad, mem_id=0, need_dx=True):
        X = self._mem[mem_id]
        ups = 1
        if isinstance(X, UpsampledInput):        # saved at low resolution; the conv saw its x2 upsampling
            X, ups = X.array, X.factor
        n, h, w, c = X.shape
        in_shape = (n, h * ups, w * ups, c)
        desc = self._desc(X.shape, ups)
        assert grad.shape == self.get_output_shapes(in_shape)[0], (
            f'{grad.shape} != {self.get_output_shapes(in_shape)[0]}')
        need = ctypes.c_size_t(0)
        lib.uocr_conv2d_wgrad_workspace(ctypes.byref(desc), ctypes.byref(need))
        ws = DeviceArray(((need.value + 3) // 4,)) if need.value else None
        lib.uocr_conv2d_wgrad(ctypes.byref(desc), X.ptr, grad.ptr, self.w.grad.ptr, self.b.grad.ptr,
                              1, ws.ptr if ws is not None else None, need.value, stream())
        if not need_dx:
            return None
        dx = DeviceArray(in_shape)               # gradient w.r.t. the (upsampled) tensor the conv read
        lib.uocr_conv2d_dgrad(ctypes.byref(self._desc(in_shape)), grad.ptr, self.w.value.ptr, dx.ptr, stream())
        return dx

    def backward_params_only(self, grads):
        """dW / db only (no dX): used for first layers when the caller does not want input gradients."""
        tracker = self.progress_tracker
        tracker.start_tracking(self.name, 'backward')
        for mem_id, grad in enumerate(make_list_if_not(grads)):
            self._backward(as_device(grad), mem_id, need_dx=False)
        self.clear_memory()
        tracker.stop_tracking(self.name, 'backward')

    def get_output_shapes(self, input_shapes):
        batch, height, width, _ = make_list_if_not(input_shapes)[0]
        (kh, kw), (ph, pw), (sh, sw) = self.kernel_size, self.padding, self.stride
        return [(batch, math.floor((height + 2 * ph - kh) / sh + 1),
                 math.floor((width + 2 * pw - kw) / sw + 1), self.out_channels)]

    def changes_receptive_field(self):
        return True

    def _get_receptive_field(self, axis, position, output_id):
        return self._window_receptive_field(axis, position, output_id, 'Convolutional2D')

    def params(self):
        return {'w': self.w, 'b': self.b}


class Conv2DToBatchedFixedWidthed(BaseLayer):
    """Sliding width-`width` windows of the W axis -> batch rows
    (layers/convolutional.py:330-373): (N, H, W, C) -> (N*W, H, width, C)."""

    def __init__(self, width, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.width = width

    def _forward(self, X, mem_id=0):
        n, h, w, c = X.shape
        y = DeviceArray(self.get_output_shapes(X.shape)[0])
        lib.uocr_window_batch_fwd(X.ptr, y.ptr, n, h, w, c, self.width, stream())
        self._mem[mem_id] = X.shape
        return y

    def _backward(self, grad, mem_id=0):
        n, h, w, c = self._mem[mem_id]
        dx = DeviceArray((n, h, w, c))
        lib.uocr_window_batch_bwd(grad.ptr, dx.ptr, n, h, w, c, self.width, stream())
        return dx

    def get_output_shapes(self, input_shapes):
        out = []
        for bs, h, w, ch in make_list_if_not(input_shapes):
            assert w >= self.width, (
                f'Input width must be >= than output width, found: {w} < {self.width}')
            out.append((bs * w, h, self.width, ch))
        return out


# ------------------------------------------------------------------------------------------
# MaxPool2D / Upsample2D
# ------------------------------------------------------------------------------------------

class MaxPool2D(BaseLayer):
    """Max pooling with zero padding, stride, ceil_mode and tie-aware backward
    (layers/maxpool.py:10-239; the CPU path :24-90 defines the semantics)."""

    def __init__(self, kernel_size, padding=0, stride=None, ceil_mode=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.kernel_size = tuplize('kernel_size', kernel_size, 2)
        self.padding = tuplize('padding', padding, 2)
        self.stride = self.kernel_size if stride is None else tuplize('stride', stride, 2)
        self.ceil_mode = ceil_mode

    def _forward(self, X, mem_id=0):
        n, h, w, c = X.shape
        out_shape = self.get_output_shapes(X.shape)[0]
        _, ho, wo, _ = out_shape
        (kh, kw) = self.kernel_size
        y = DeviceArray(out_shape)
        mask = DeviceArray((n, kh * ho, kw * wo, c), np.uint8)
        lib.uocr_maxpool2d_fwd(X.ptr, y.ptr, mask.ptr, n, h, w, c, kh, kw, *self.padding,
                               *self.stride, ho, wo, stream())
        self._mem[mem_id] = mask, X.shape
        return y

    def _backward(self, grad, mem_id=0):
        mask, (n, h, w, c) = self._mem[mem_id]
        _, ho, wo, _ = grad.shape
        dx = DeviceArray((n, h, w, c))
        lib.uocr_maxpool2d_bwd(grad.ptr, mask.ptr, dx.ptr, n, h, w, c, *self.kernel_size,
                               *self.padding, *self.stride, ho, wo, stream())
        return dx

    def get_output_shapes(self, input_shapes):
        batch, height, width, channels = make_list_if_not(input_shapes)[0]
        (kh, kw), (ph, pw), (sh, sw) = self.kernel_size, self.padding, self.stride
        rnd = math.ceil if self.ceil_mode else math.floor
        return [(batch, rnd((height + 2 * ph - kh) / sh + 1), rnd((width + 2 * pw - kw) / sw + 1),
                 channels)]

    def changes_receptive_field(self):
        return True

    def _get_receptive_field(self, axis, position, output_id):
        return self._window_receptive_field(axis, position, output_id, 'MaxPool2D')


class Upsample2D(BaseLayer):
    """Nearest-neighbour repeat by (sy, sx); backward = block sum (layers/upsample.py:10-135)."""

    def __init__(self, scale_factor, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.scale_factor = tuplize('scale_factor', scale_factor, 2)

    def _forward(self, X, mem_id=0):
        n, h, w, c = X.shape
        self._mem[mem_id] = X.shape
        y = DeviceArray(self.get_output_shapes(X.shape)[0])
        lib.uocr_upsample2d_fwd(X.ptr, y.ptr, n, h, w, c, *self.scale_factor, stream())
        return y

    def _backward(self, grad, mem_id=0):
        n, h, w, c = self._mem[mem_id]
        dx = DeviceArray((n, h, w, c))
        lib.uocr_upsample2d_bwd(grad.ptr, dx.ptr, n, h, w, c, *self.scale_factor, stream())
        return dx

    def get_output_shapes(self, input_shapes):
        n, h, w, c = make_list_if_not(input_shapes)[0]
        return [(n, h * self.scale_factor[0], w * self.scale_factor[1], c)]

    def changes_receptive_field(self):
        return True

    def _get_receptive_field(self, axis, position, output_id):
        assert 0 <= axis < 2, f'Upsample2D has two axis, found {axis}'
        assert output_id < self.get_outputs_count(), (
            f'This layer has only {self.get_outputs_count()} outputs')
        key = (axis, position, output_id)
        if key not in self._receptive_fields:
            self._receptive_fields[key] = {0: {position // self.scale_factor[axis]}}
        return self._receptive_fields[key]
