import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import glue, my_model
from univer_ocr_b200._lib import lib
nn.CP.use_gpu()
B = 64
mono = my_model.make_monochrome((B, 496, 736, 1)); para = my_model.make_paragraph((B, 496, 736, 1))
u8 = nn.CP.pinned_empty((B, 496, 736, 1), np.uint8); u8[...] = 200
res = ctypes = None
import ctypes
def reserved():
    n = ctypes.c_size_t(0); lib.uocr_mempool_reserved(ctypes.byref(n)); return n.value / 1e6
out_host = None
for it in range(8):
    t = [time.perf_counter()]
    d = nn.DeviceArray.from_host(u8, np.uint8); t.append(time.perf_counter())
    x = glue.pixels_to_unit(d); t.append(time.perf_counter())
    m = mono.predict(x)[0]; t.append(time.perf_counter())
    p = para.predict(m)[0]; t.append(time.perf_counter())
    k = glue.thresholded(p); t.append(time.perf_counter())
    if out_host is None: out_host = nn.CP.pinned_empty(k.shape, np.uint8)
    lib.uocr_memcpy_d2h(out_host.ctypes.data, k.ptr, k.nbytes, nn.CP.stream()); t.append(time.perf_counter())
    nn.CP.synchronize(); t.append(time.perf_counter())
    print(it, ' '.join(f'{(b - a) * 1e3:.2f}' for a, b in zip(t, t[1:])), f'reserved {reserved():.0f} MB', flush=True)
