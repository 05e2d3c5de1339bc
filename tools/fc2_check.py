"""uocr_fc_chain2_fwd (dense_2 + LeakyRelu + dense_3 in one kernel) against float64 NumPy, and its time next to the two
separate FullyConnected calls."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200._lib import lib, MATH_TF32, ACT_LEAKY, ACT_NONE

nn.CP.use_gpu(); nn.CP.set_math_mode('tf32')
rng = np.random.default_rng(0)
st = nn.CP.stream()


def ev():
    e = ctypes.c_void_p(); lib.uocr_event_create(ctypes.byref(e)); return e.value


def timed(fn, reps=50):
    for _ in range(5): fn()
    e0, e1 = ev(), ev()
    lib.uocr_event_record(e0, st)
    for _ in range(reps): fn()
    lib.uocr_event_record(e1, st); lib.uocr_event_sync(e1)
    ms = ctypes.c_float(0); lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    return ms.value / reps * 1e3


worst = 0.0
for (M, K1, N2) in ((16384, 1024, 162), (1000, 1024, 162), (4096, 64, 16), (640, 256, 256), (777, 128, 200)):
    X = nn.CP.copy(rng.standard_normal((M, K1)).astype(np.float32))
    W1 = nn.CP.copy((rng.standard_normal((K1 + 1, 128)) / np.sqrt(K1)).astype(np.float32))
    W2 = nn.CP.copy((rng.standard_normal((129, N2)) / np.sqrt(128)).astype(np.float32))
    w1t = nn.DeviceArray((128, K1)); lib.uocr_weights_to_kmajor(W1.ptr, w1t.ptr, K1, 128, st)
    w2t = nn.DeviceArray((N2, 128)); lib.uocr_weights_to_kmajor(W2.ptr, w2t.ptr, 128, N2, st)
    y = nn.DeviceArray.full((M, N2), -7.0)
    hid = nn.DeviceArray((M, 128)); y2 = nn.DeviceArray((M, N2))

    def fused():
        lib.uocr_fc_chain2_fwd(X.ptr, W1.ptr, w1t.ptr, W2.ptr, w2t.ptr, y.ptr, M, K1, 128, N2, ACT_LEAKY, 0.01, MATH_TF32, st)

    def separate():
        lib.uocr_fc_fwd_kmajor(X.ptr, W1.ptr, w1t.ptr, hid.ptr, M, K1, 128, ACT_LEAKY, 0.01, MATH_TF32, st)
        lib.uocr_fc_fwd_kmajor(hid.ptr, W2.ptr, w2t.ptr, y2.ptr, M, 128, N2, ACT_NONE, 0.0, MATH_TF32, st)
    fused(); separate()
    x64, a, b = X.get().astype(np.float64), W1.get().astype(np.float64), W2.get().astype(np.float64)
    h = x64 @ a[:-1] + a[-1]
    h = np.where(h > 0, h, 0.01 * h)
    want = h @ b[:-1] + b[-1]
    got, got2 = y.get().astype(np.float64), y2.get().astype(np.float64)
    err, err2 = np.abs(got - want).max() / np.abs(want).max(), np.abs(got2 - want).max() / np.abs(want).max()
    worst = max(worst, err)
    print(f'M={M} K1={K1} N2={N2}: fused err {err:.1e} (separate {err2:.1e})  fused {timed(fused):.1f} us  separate {timed(separate):.1f} us', flush=True)
# without cached K-major copies
lib.uocr_fc_chain2_fwd(X.ptr, W1.ptr, None, W2.ptr, None, y.ptr, M, K1, 128, N2, ACT_LEAKY, 0.01, MATH_TF32, st)
print('no k-major copies: err', np.abs(y.get() - want).max() / np.abs(want).max())
sys.exit(0 if worst < 2e-3 else 1)
