"""Times uocr_fc_fwd_kmajor (persistent tcgen05 GEMM) on a few shapes; UOCR_TC_PAIR selects the CTA-pair kernel."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200._lib import lib, MATH_TF32, ACT_LEAKY

nn.CP.use_gpu(); nn.CP.set_math_mode('tf32')
rng = np.random.default_rng(0)


def ev():
    e = ctypes.c_void_p(); lib.uocr_event_create(ctypes.byref(e)); return e.value


SHAPES = [(16384, 512, 1024), (16384, 1024, 512), (65536, 1024, 1024), (16384, 1024, 128)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in sys.argv[1].split('x'))]
for (M, K, N) in SHAPES:
    X = nn.CP.copy(rng.standard_normal((M, K)).astype(np.float32))
    W = nn.CP.copy((rng.standard_normal((K + 1, N)) / np.sqrt(K)).astype(np.float32))
    wt = nn.DeviceArray((N, K)); lib.uocr_weights_to_kmajor(W.ptr, wt.ptr, K, N, nn.CP.stream())
    y = nn.DeviceArray((M, N))
    st = nn.CP.stream()
    def run():
        lib.uocr_fc_fwd_kmajor(X.ptr, W.ptr, wt.ptr, y.ptr, M, K, N, ACT_LEAKY, 0.01, MATH_TF32, st)
    for _ in range(5): run()
    e0, e1 = ev(), ev()
    reps = 50
    lib.uocr_event_record(e0, st)
    for _ in range(reps): run()
    lib.uocr_event_record(e1, st); lib.uocr_event_sync(e1)
    ms = ctypes.c_float(0); lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    t = ms.value / reps
    ref = np.asarray(X.get()[:64], dtype=np.float64) @ np.asarray(W.get()[:-1], dtype=np.float64) + np.asarray(W.get()[-1], dtype=np.float64)
    ref = np.where(ref > 0, ref, 0.01 * ref)
    err = np.max(np.abs(y.get()[:64] - ref)) / np.max(np.abs(ref))
    print(f'M={M} K={K} N={N}: {t*1e3:.1f} us  {2*M*K*N/t/1e9:.0f} TFLOP/s  err {err:.1e}', flush=True)
