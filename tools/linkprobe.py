"""Host-link probe: pinned H2D / D2H bandwidth alone and both directions at once, at the e2e step's payload sizes."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200._lib import lib
from univer_ocr_b200.pipeline import _new_stream

nn.CP.use_gpu()
up, down = 25985024, 30212096
h_in, h_out = nn.CP.pinned_empty((up,), np.uint8), nn.CP.pinned_empty((down,), np.uint8)
d_in, d_out = nn.DeviceArray.empty((up,), np.uint8), nn.DeviceArray.empty((down,), np.uint8)
s1, s2 = _new_stream(), _new_stream()


def run(do_up, do_down, reps=200):
    nn.CP.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if do_up:
            lib.uocr_memcpy_h2d(d_in.ptr, h_in.ctypes.data, up, s1)
        if do_down:
            lib.uocr_memcpy_d2h(h_out.ctypes.data, d_out.ptr, down, s2)
    lib.uocr_stream_sync(s1); lib.uocr_stream_sync(s2)
    dt = (time.perf_counter() - t0) / reps
    return dt


for name, a, b in (('h2d alone', 1, 0), ('d2h alone', 0, 1), ('both', 1, 1)):
    run(a, b, 20)
    dt = run(a, b)
    print(f'{name}: {dt * 1e3:.3f} ms per step  up {a * up / dt / 1e9:.1f} GB/s  down {b * down / dt / 1e9:.1f} GB/s  -> {64 / dt:.0f} images/s ceiling', flush=True)
