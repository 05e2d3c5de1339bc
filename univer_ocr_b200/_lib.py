"""ctypes binding of libuocr.so -- the thin shim north_star asks for: Python passes raw device
pointers, shapes and a stream; nothing here computes.

The argument types are derived from `include/uocr.h` itself, so the binding cannot drift from
the C ABI.  There is NO fallback: if the shared library is missing (`python -m
univer_ocr_b200.build` was not run) or no CUDA device is visible, the first compute call raises.
"""
import ctypes
import os
import re

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
HEADER = os.path.join(ROOT, 'include', 'uocr.h')
LIB_PATH = os.path.join(PKG_DIR, 'lib', 'libuocr.so')

UOCR_OK = 0
ACT_NONE, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2
MATH_FP32, MATH_TF32 = 0, 1
SEG_DICE, SEG_JACCARD = 0, 1
REG_L1, REG_L2 = 1, 2


class UocrError(RuntimeError):
    def __init__(self, code, func, message):
        super().__init__(f'{func} failed with code {code}: {message}')
        self.code = code


class ConvDesc(ctypes.Structure):
    """struct uocr_conv2d_desc (include/uocr.h)."""
    _fields_ = [('n', ctypes.c_int64), ('h', ctypes.c_int64), ('w', ctypes.c_int64),
                ('cin', ctypes.c_int64), ('cout', ctypes.c_int64),
                ('kh', ctypes.c_int32), ('kw', ctypes.c_int32),
                ('ph', ctypes.c_int32), ('pw', ctypes.c_int32),
                ('sh', ctypes.c_int32), ('sw', ctypes.c_int32),
                ('padding_value', ctypes.c_float), ('bias', ctypes.c_int32),
                ('math_mode', ctypes.c_int32), ('in_upsample', ctypes.c_int32)]


_SCALARS = {
    'int': ctypes.c_int, 'int32_t': ctypes.c_int32, 'int64_t': ctypes.c_int64,
    'uint64_t': ctypes.c_uint64, 'size_t': ctypes.c_size_t, 'float': ctypes.c_float,
}

_DECL = re.compile(r'^\s*(int|const char\*)\s+(uocr_\w+)\s*\(([^;{]*?)\)\s*;', re.M | re.S)


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every `uocr_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for ret, name, args in _DECL.findall(text):
        argtypes = []
        args = ' '.join(args.split())
        if args and args != 'void':
            for arg in args.split(','):
                arg = arg.strip()
                if '*' in arg:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ctype = arg.replace('const ', '').split()[0]
                    argtypes.append(_SCALARS[ctype])
        protos[name] = (ctypes.c_char_p if 'char' in ret else ctypes.c_int, argtypes)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self._protos = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise UocrError(-100, 'load', f'{LIB_PATH} not found -- build it with '
                            f'`python -m univer_ocr_b200.build` (there is no CPU fallback)')
        dll = ctypes.CDLL(LIB_PATH)
        self._protos = parse_header()
        for name, (restype, argtypes) in self._protos.items():
            fn = getattr(dll, name)          # AttributeError if the library misses a symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def __getattr__(self, name):
        if not name.startswith('uocr_'):
            raise AttributeError(name)
        dll = self.load()
        fn = getattr(dll, name)
        if self._protos[name][0] is not ctypes.c_int or name == 'uocr_version':
            return fn

        def checked(*args):
            rc = fn(*args)
            if rc != UOCR_OK:
                raise UocrError(rc, name, dll.uocr_last_error().decode(errors='replace'))
            return rc
        checked.__name__ = name
        setattr(self, name, checked)          # cache: next lookup bypasses __getattr__
        return checked

    def exported_names(self):
        self.load()
        return sorted(self._protos)


lib = _Lib()


def device_count():
    n = ctypes.c_int(0)
    try:
        lib.uocr_device_count(ctypes.byref(n))
    except UocrError:
        return 0
    return n.value


def require_device():
    if device_count() < 1:
        raise UocrError(-2, 'require_device',
                        'no CUDA device visible: univer_ocr_b200 has no CPU path')


def launch_count():
    n = ctypes.c_uint64(0)
    lib.uocr_launch_count(ctypes.byref(n))
    return n.value
