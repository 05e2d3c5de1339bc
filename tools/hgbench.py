"""Times uocr_hourglass1_fwd_mode (Paragraph network in one kernel) at batch 64 in FP32 and TF32 mode."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200._lib import ACT_SIGMOID, lib
nn.CP.use_gpu()
rng = np.random.default_rng(1)
n, h, w = 64, 496, 736
X = nn.CP.copy(rng.uniform(size=(n, h, w, 1)).astype(np.float32))
ws = [nn.CP.copy((rng.standard_normal((5, 5, 1, 1)) * 0.25).astype(np.float32)) for _ in range(5)]
bs = [nn.CP.copy((rng.standard_normal(1) * 0.3).astype(np.float32)) for _ in range(5)]
ptrs = ctypes.c_void_p * 5
wp, bp = ptrs(*[a.ptr for a in ws]), ptrs(*[a.ptr for a in bs])
y = nn.DeviceArray((n, h, w, 1))
st = nn.CP.stream()
flush = nn.DeviceArray((64 * 1024 * 1024,))
def ev():
    e = ctypes.c_void_p(); lib.uocr_event_create(ctypes.byref(e)); return e.value
ref = None
for mode in (0, 1):
    ts = []
    for i in range(8):
        flush.fill(0)
        e0, e1 = ev(), ev()
        lib.uocr_event_record(e0, st)
        lib.uocr_hourglass1_fwd_mode(X.ptr, wp, bp, y.ptr, n, h, w, 0.01, ACT_SIGMOID, 0.0, mode, st)
        lib.uocr_event_record(e1, st); lib.uocr_event_sync(e1)
        ms = ctypes.c_float(0); lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms)); ts.append(ms.value)
    out = y.get().astype(np.float64)
    if ref is None: ref = out
    t = float(np.median(ts[2:]))
    print(f'mode {mode}: {t * 1e3:.1f} us per 64 tiles ({186.9e6 / t / 1e6 / 6547.8 * 100:.1f} % of HBM peak), max |diff to fp32| {np.max(np.abs(out - ref)):.2e}', flush=True)
