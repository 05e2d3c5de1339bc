"""Array backend of the B200 path: the counterpart of the reference's `nn/gpu.py` (:5-29).

The reference flips `CP.cp` between numpy and cupy.  Here there is exactly one backend --
device memory owned by libuocr (stream-ordered pool) wrapped in `DeviceArray` -- and NO CPU
path: `CP.use_cpu()` raises.  CuPy is not required (it is absent from this image); any object
exposing `__cuda_array_interface__` (a CuPy array, a torch CUDA tensor) is accepted as input
and `DeviceArray` exposes the same protocol, so the two interoperate zero-copy.

Storage is float32 (the reference is float64); see DESIGN.md for the tolerance contract.
"""
import contextlib
import ctypes
import os
import weakref

import numpy as np

from .._lib import MATH_FP32, MATH_TF32, lib, require_device

_F32 = np.dtype(np.float32)


class _Runtime:
    """Process-wide device + stream (one process per GPU)."""

    def __init__(self):
        self.device = None
        self.stream = None          # int (cudaStream_t) or None before init

    def ensure(self):
        if self.stream is not None:
            return self
        require_device()
        dev = int(os.environ.get('UOCR_DEVICE', os.environ.get('LOCAL_RANK', '0')))
        n = ctypes.c_int(0)
        lib.uocr_device_count(ctypes.byref(n))
        self.device = dev % max(n.value, 1)
        lib.uocr_set_device(self.device)
        _bind_to_gpu_numa_node(self.device)
        s = ctypes.c_void_p()
        lib.uocr_stream_create(ctypes.byref(s))
        self.stream = s.value
        return self


def _bind_to_gpu_numa_node(device):
    """Pin this process to the CPUs NVML reports as local to the GPU, so that pinned host buffers (first touched by
    cudaHostAlloc on this thread) land on the GPU's NUMA node: with one process per GPU on a two-socket host the
    H2D / D2H copies of the end-to-end path otherwise cross the socket interconnect.  Best effort; UOCR_NO_AFFINITY=1
    disables it."""
    if os.environ.get('UOCR_NO_AFFINITY') == '1' or not hasattr(os, 'sched_setaffinity'):
        return
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        index = device
        if visible:
            ids = [v.strip() for v in visible.split(',') if v.strip()]
            if device < len(ids) and ids[device].isdigit():
                index = int(ids[device])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * i + b for i, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:                                     # no NVML / not permitted: keep the inherited affinity
        pass


RT = _Runtime()


def stream():
    return RT.ensure().stream


class _Buffer:
    """Owns one pooled device allocation; freed (stream-ordered) when the last view dies."""
    __slots__ = ('ptr', 'nbytes', '__weakref__')

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        st = stream()
        lib.uocr_malloc(ctypes.byref(p), max(int(nbytes), 1), st)
        self.ptr = p.value
        self.nbytes = int(nbytes)
        weakref.finalize(self, _free, self.ptr, st)


def _free(ptr, st):
    # back to the pool of the stream it was allocated on (reuse is ordered on THAT stream): with several compute
    # streams (CP.on_stream) a block must not migrate to another stream's free list while its own stream may still
    # have kernels in flight on it
    try:
        if RT.stream is not None:
            lib.uocr_free(ptr, st)
    except Exception:       # interpreter shutdown
        pass


class _Foreign:
    """Keeps a foreign __cuda_array_interface__ owner alive."""
    __slots__ = ('ptr', 'nbytes', 'owner')

    def __init__(self, ptr, nbytes, owner):
        self.ptr, self.nbytes, self.owner = ptr, nbytes, owner


def _prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class DeviceArray:
    """Dense C-contiguous device tensor (float32 unless stated).  Duck-compatible with the
    subset of numpy/cupy ndarray the reference's layer stack and graph executor use:
    shape/size/ndim/dtype, reshape, copy, tolist, `+`, `+=`, `0 + a` (models.py:218 sums
    fan-out gradients with Python's sum()), scalar `*` and `/`."""
    __array_priority__ = 100.0

    def __init__(self, shape, dtype=_F32, buf=None, offset=0):
        if isinstance(shape, (int, np.integer)):
            shape = (shape,)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = _prod(self.shape)
        self.nbytes = self.size * self.dtype.itemsize
        self._buf = buf if buf is not None else _Buffer(self.nbytes)
        self.ptr = self._buf.ptr + int(offset)

    # ---- construction -------------------------------------------------------------
    @staticmethod
    def empty(shape, dtype=_F32):
        return DeviceArray(shape, dtype)

    @staticmethod
    def zeros(shape, dtype=_F32):
        a = DeviceArray(shape, dtype)
        lib.uocr_memset(a.ptr, 0, a.nbytes, stream())
        return a

    @staticmethod
    def full(shape, value):
        a = DeviceArray(shape)
        lib.uocr_fill_f32(a.ptr, float(value), a.size, stream())
        return a

    @staticmethod
    def from_host(obj, dtype=_F32):
        """H2D copy (CP.copy, gpu.py:18-22).  Host data are converted to `dtype` on the host."""
        host = np.ascontiguousarray(obj, dtype=dtype)
        a = DeviceArray(host.shape, dtype)
        lib.uocr_memcpy_h2d(a.ptr, host.ctypes.data, host.nbytes, stream())
        if not getattr(obj, '_uocr_pinned', False):
            lib.uocr_stream_sync(stream())      # pageable source may be freed by the caller
        return a

    @staticmethod
    def from_cuda_array_interface(obj):
        cai = obj.__cuda_array_interface__
        if cai.get('strides') is not None:
            raise ValueError('only C-contiguous device arrays are supported')
        dt = np.dtype(cai['typestr'])
        shape = tuple(cai['shape'])
        return DeviceArray(shape, dt, buf=_Foreign(cai['data'][0], _prod(shape) * dt.itemsize, obj))

    # ---- protocol -----------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def __cuda_array_interface__(self):
        return {'shape': self.shape, 'typestr': self.dtype.str, 'data': (self.ptr, False),
                'version': 3, 'strides': None}

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f'DeviceArray(shape={self.shape}, dtype={self.dtype.name})'

    # ---- host transfer ------------------------------------------------------------
    def get(self, out=None):
        """D2H copy (CP.asnumpy, gpu.py:24-28); synchronises the stream."""
        host = out if out is not None else np.empty(self.shape, dtype=self.dtype)
        lib.uocr_memcpy_d2h(host.ctypes.data, self.ptr, self.nbytes, stream())
        lib.uocr_stream_sync(stream())
        return host

    def __array__(self, dtype=None, copy=None):
        host = self.get()
        return host if dtype is None else host.astype(dtype)

    def tolist(self):
        return self.get().astype(np.float64).tolist()

    def item(self):
        assert self.size == 1
        return float(self.get().reshape(-1)[0])

    def __float__(self):
        return self.item()

    # ---- views / copies -----------------------------------------------------------
    def reshape(self, *shape):
        if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)):
            shape = tuple(shape[0])
        shape = [int(s) for s in shape]
        if -1 in shape:
            known = _prod(s for s in shape if s != -1)
            shape[shape.index(-1)] = self.size // max(known, 1)
        assert _prod(shape) == self.size, f'cannot reshape {self.shape} into {tuple(shape)}'
        return DeviceArray(shape, self.dtype, buf=self._buf, offset=self.ptr - self._buf.ptr)

    def ravel(self):
        return self.reshape(self.size)

    def copy(self):
        out = DeviceArray(self.shape, self.dtype)
        lib.uocr_memcpy_d2d(out.ptr, self.ptr, self.nbytes, stream())
        return out

    def flat_view(self, start, count, shape=None):
        """Contiguous sub-range [start, start+count) of the flattened array, as a view."""
        assert 0 <= start and start + count <= self.size
        return DeviceArray(shape if shape is not None else (count,), self.dtype, buf=self._buf,
                           offset=self.ptr - self._buf.ptr + start * self.dtype.itemsize)

    def __getitem__(self, idx):
        # leading-axis integer / contiguous slice -> view; anything else -> host round trip
        if isinstance(idx, (int, np.integer)):
            i = int(idx) % self.shape[0]
            inner = _prod(self.shape[1:])
            return self.flat_view(i * inner, inner, self.shape[1:])
        if isinstance(idx, slice) and idx.step in (None, 1):
            lo, hi, _ = idx.indices(self.shape[0])
            inner = _prod(self.shape[1:])
            return self.flat_view(lo * inner, max(hi - lo, 0) * inner,
                                  (max(hi - lo, 0),) + self.shape[1:])
        return self.get()[idx]

    def fill(self, value):
        if value == 0:
            lib.uocr_memset(self.ptr, 0, self.nbytes, stream())
        else:
            assert self.dtype == _F32
            lib.uocr_fill_f32(self.ptr, float(value), self.size, stream())
        return self

    def set(self, obj):
        """In-place H2D upload."""
        host = np.ascontiguousarray(obj, dtype=self.dtype)
        assert host.shape == self.shape, f'{host.shape} != {self.shape}'
        lib.uocr_memcpy_h2d(self.ptr, host.ctypes.data, host.nbytes, stream())
        lib.uocr_stream_sync(stream())
        return self

    # ---- arithmetic (float32 only) ------------------------------------------------
    def _axpby(self, x, a, b):
        lib.uocr_axpby_f32(self.ptr, x.ptr, float(a), float(b), self.size, stream())
        return self

    def _coerce(self, other):
        if isinstance(other, DeviceArray):
            assert other.shape == self.shape, f'shape mismatch {self.shape} vs {other.shape}'
            return other
        if isinstance(other, np.ndarray) and other.shape == self.shape:
            return DeviceArray.from_host(other)
        return None

    def __add__(self, other):
        o = self._coerce(other)
        if o is not None:
            return self.copy()._axpby(o, 1.0, 1.0)
        if np.isscalar(other):
            if other == 0:
                return self              # `sum()` starts from int 0 (models.py:218)
            return DeviceArray.full(self.shape, other)._axpby(self, 1.0, 1.0)
        return NotImplemented

    __radd__ = __add__

    def __iadd__(self, other):
        o = self._coerce(other)
        if o is not None:
            return self._axpby(o, 1.0, 1.0)
        if np.isscalar(other):
            return self._axpby(DeviceArray.full(self.shape, other), 1.0, 1.0)
        return NotImplemented

    def __sub__(self, other):
        o = self._coerce(other)
        if o is not None:
            return self.copy()._axpby(o, -1.0, 1.0)
        if np.isscalar(other):
            return self + (-other)
        return NotImplemented

    def __isub__(self, other):
        o = self._coerce(other)
        if o is not None:
            return self._axpby(o, -1.0, 1.0)
        return NotImplemented

    def __neg__(self):
        return DeviceArray(self.shape)._axpby(self, -1.0, 0.0)

    def __mul__(self, other):
        if np.isscalar(other):
            return DeviceArray(self.shape)._axpby(self, other, 0.0)
        o = self._coerce(other)
        if o is not None:
            out = DeviceArray(self.shape)
            lib.uocr_mul_f32(out.ptr, self.ptr, o.ptr, self.size, stream())
            return out
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, other):
        if np.isscalar(other):
            return self * (1.0 / other)
        return NotImplemented

    def sum(self):
        out = DeviceArray((1,))
        lib.uocr_sum_f32(self.ptr, self.size, out.ptr, 0, stream())
        return out.item()

    def isnan_any(self):
        flag = DeviceArray.zeros((1,), np.int32)
        lib.uocr_nan_flag_f32(self.ptr, self.size, flag.ptr, stream())
        return bool(flag.get()[0])


class LazyScalar:
    """A float that still lives on the device.  The reference turns every loss into a Python
    float at once (`float(loss)`, losses.py:25,73; regularizations.py:19,26) -- one forced device
    sync per call.  This object defers the read-back until the value is actually used; any
    arithmetic or formatting fetches it (once) and behaves like a float."""
    __slots__ = ('_dev', '_val')

    def __init__(self, dev):
        self._dev = dev
        self._val = None

    def __float__(self):
        if self._val is None:
            self._val = self._dev.item()
            self._dev = None
        return self._val

    def peek(self):
        """The value on the device right now, without caching it: a step replayed from a CUDA graph hands back the
        same LazyScalar objects every time, over buffers the replay rewrites."""
        return self._dev.item() if self._dev is not None else self._val

    def _f(op):
        def fn(self, other):
            return getattr(float(self), op)(float(other))
        return fn

    __add__, __radd__, __sub__, __rsub__ = _f('__add__'), _f('__radd__'), _f('__sub__'), _f('__rsub__')
    __mul__, __rmul__ = _f('__mul__'), _f('__rmul__')
    __truediv__, __rtruediv__ = _f('__truediv__'), _f('__rtruediv__')
    __lt__, __le__, __gt__, __ge__ = _f('__lt__'), _f('__le__'), _f('__gt__'), _f('__ge__')
    __eq__, __ne__ = _f('__eq__'), _f('__ne__')
    __hash__ = None
    del _f

    def __neg__(self):
        return -float(self)

    def __abs__(self):
        return abs(float(self))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype)

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return repr(float(self))


class _DeviceNamespace:
    """What `CP.cp` resolves to: the handful of numpy-style constructors the layer stack uses."""
    ndarray = DeviceArray
    float32 = np.float32

    @staticmethod
    def array(obj, dtype=np.float32):
        return CP.copy(obj) if not isinstance(obj, DeviceArray) else obj.copy()

    asarray = array

    @staticmethod
    def zeros(shape, dtype=np.float32):
        return DeviceArray.zeros(shape if not isinstance(shape, (int, np.integer)) else (shape,), dtype)

    @staticmethod
    def ones(shape, dtype=np.float32):
        return DeviceArray.full(shape if not isinstance(shape, (int, np.integer)) else (shape,), 1.0)

    @staticmethod
    def empty(shape, dtype=np.float32):
        return DeviceArray.empty(shape if not isinstance(shape, (int, np.integer)) else (shape,), dtype)

    @staticmethod
    def zeros_like(a):
        return DeviceArray.zeros(a.shape, getattr(a, 'dtype', np.float32)
                                 if isinstance(a, DeviceArray) else np.float32)

    @staticmethod
    def ones_like(a):
        return DeviceArray.full(a.shape, 1.0)

    @staticmethod
    def copy(a):
        return a.copy() if isinstance(a, DeviceArray) else CP.copy(a)

    @staticmethod
    def reshape(a, shape):
        return a.reshape(shape)

    @staticmethod
    def concatenate(arrays, axis=-1):
        return concatenate(arrays, axis)


def concatenate(arrays, axis=-1):
    """Device-side np.concatenate along `axis` (Concat.forward, layers.py:252)."""
    arrays = [CP.copy(a) if not isinstance(a, DeviceArray) else a for a in arrays]
    nd = arrays[0].ndim
    axis = axis % nd
    outer = _prod(arrays[0].shape[:axis])
    inner = [_prod(a.shape[axis:]) for a in arrays]
    shape = list(arrays[0].shape)
    shape[axis] = sum(a.shape[axis] for a in arrays)
    out = DeviceArray(shape)
    pitch, pos = sum(inner), 0
    for a, cols in zip(arrays, inner):
        lib.uocr_copy2d_f32(out.ptr + 4 * pos, pitch, a.ptr, cols, outer, cols, stream())
        pos += cols
    return out


def slice_axis(a, axis, lo, hi):
    """Device-side copy of `a[..., lo:hi, ...]` along `axis` (Concat.backward, layers.py:262-266)."""
    axis = axis % a.ndim
    outer = _prod(a.shape[:axis])
    tail = _prod(a.shape[axis + 1:])
    shape = list(a.shape)
    shape[axis] = hi - lo
    out = DeviceArray(shape)
    lib.uocr_copy2d_f32(out.ptr, (hi - lo) * tail, a.ptr + 4 * lo * tail, a.shape[axis] * tail,
                        outer, (hi - lo) * tail, stream())
    return out


class CP:
    """Same surface as the reference's `CP` (nn/gpu.py:5-29)."""
    cp = _DeviceNamespace
    is_gpu_used = True
    # TF32 (tcgen05 tensor-core kernels, FP32 accumulate, 1e-3 tolerance contract) is the default on sm_100 devices:
    # a drop-in user gets the fast path without knowing about `set_math_mode`.  `set_math_mode('fp32')` (or
    # UOCR_MATH=fp32) selects the FP32-FFMA check mode; the first `use_gpu()` settles the default (`_math_mode_set`).
    math_mode = MATH_TF32
    _math_mode_set = False
    # bumped by everything that changes parameter values (Param.value = ..., optimizer updates, the flat-buffer
    # DataParallel update): layers key their cached K-major weight copies on it
    weights_generation = 0

    @staticmethod
    def use_gpu():
        RT.ensure()
        CP.is_gpu_used = True
        if not CP._math_mode_set:
            CP._math_mode_set = True
            env = os.environ.get('UOCR_MATH')
            if env in ('fp32', 'tf32'):
                CP.math_mode = {'fp32': MATH_FP32, 'tf32': MATH_TF32}[env]
            else:
                major = ctypes.c_int(0)
                lib.uocr_device_info(RT.device, None, 0, None, ctypes.byref(major), None, None, None)
                CP.math_mode = MATH_TF32 if major.value == 10 else MATH_FP32     # tcgen05 kernels are sm_100a only

    @staticmethod
    def use_cpu():
        raise RuntimeError('univer_ocr_b200 is the B200 path of the layer stack: it has no CPU '
                           'fallback (use the reference package for CPU runs)')

    @staticmethod
    def set_math_mode(mode):
        """'fp32' -- FFMA everywhere (check mode); 'tf32' -- tcgen05 TF32 tensor-core kernels for
        the dense contractions (Char 64->64 convolutions, FullyConnected), FP32 accumulate."""
        CP.math_mode = {'fp32': MATH_FP32, 'tf32': MATH_TF32}[mode]
        CP._math_mode_set = True

    @staticmethod
    def copy(obj):
        """Host -> device (numpy / nested lists), or a device-side copy."""
        if isinstance(obj, DeviceArray):
            return obj.copy()
        if hasattr(obj, '__cuda_array_interface__'):
            return DeviceArray.from_cuda_array_interface(obj)
        return DeviceArray.from_host(obj)

    @staticmethod
    def asnumpy(obj):
        if isinstance(obj, DeviceArray):
            return obj.get()
        if isinstance(obj, LazyScalar):
            return np.asarray(float(obj))
        return np.asarray(obj)

    @staticmethod
    def synchronize():
        lib.uocr_stream_sync(stream())

    @staticmethod
    def stream():
        return stream()

    @staticmethod
    @contextlib.contextmanager
    def on_stream(s):
        """Everything issued inside the block (kernels, allocations, copies) goes to CUDA stream `s` instead of the
        process's compute stream; ordering against other streams is the caller's business (events)."""
        RT.ensure()
        prev, RT.stream = RT.stream, s
        try:
            yield s
        finally:
            RT.stream = prev

    @staticmethod
    def pinned_empty(shape, dtype=np.float32):
        """Page-locked host array for asynchronous H2D / D2H copies."""
        dtype = np.dtype(dtype)
        nbytes = _prod(shape) * dtype.itemsize
        p = ctypes.c_void_p()
        RT.ensure()
        lib.uocr_host_alloc(ctypes.byref(p), nbytes)
        raw = (ctypes.c_byte * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(raw, dtype=dtype, count=_prod(shape)).reshape(shape).view(_Pinned)
        arr._uocr_pinned = True
        weakref.finalize(raw, lib.uocr_host_free, p.value)
        arr._uocr_raw = raw
        return arr


class _Pinned(np.ndarray):
    _uocr_pinned = True
    _uocr_raw = None


def as_device(x):
    """Layer inputs may be DeviceArrays, foreign CUDA arrays or host data (uploaded)."""
    if isinstance(x, DeviceArray):
        return x
    materialize = getattr(x, 'materialize', None)          # lazy device values (losses.LazySegGrad)
    if materialize is not None:
        return materialize()
    return CP.copy(x)
