"""Host-side time of every C-ABI call during one inference step (finds blocking calls)."""
import os, sys, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import my_model, _lib

nn.CP.use_gpu(); nn.CP.set_math_mode(sys.argv[1] if len(sys.argv) > 1 else 'tf32')
B = 64
log = []
orig = {}
for name in _lib.lib.exported_names():
    fn = getattr(_lib.lib, name)
    def make(fn, name):
        def wrapped(*a):
            t0 = time.perf_counter(); r = fn(*a); dt = time.perf_counter() - t0
            log.append((name, dt, a[1] if name == 'uocr_malloc' else None)); return r
        return wrapped
    setattr(_lib.lib, name, make(fn, name))
models = {'mono': my_model.make_monochrome((B, 496, 736, 1)), 'para': my_model.make_paragraph((B, 496, 736, 1)),
          'line': my_model.make_line((B, 128, 256, 1)), 'char': my_model.make_char((B, 32, 256, 1))}
hp = nn.CP.pinned_empty((B, 496, 736, 1)); hp[...] = 0.5
hl = nn.CP.pinned_empty((B, 128, 256, 1)); hl[...] = 0.5
hc = nn.CP.pinned_empty((B, 32, 256, 1)); hc[...] = 0.5
outs_host = None
def step():
    global outs_host
    xp, xl, xc = nn.CP.copy(hp), nn.CP.copy(hl), nn.CP.copy(hc)
    m = models['mono'].predict(xp)[0]; p = models['para'].predict(m)[0]
    l = models['line'].predict(xl)[0]; c = models['char'].predict(xc)[0]
    outs = (p, l, c)
    if outs_host is None:
        outs_host = [nn.CP.pinned_empty(o.shape) for o in outs]
    for o, h in zip(outs, outs_host):
        _lib.lib.uocr_memcpy_d2h(h.ctypes.data, o.ptr, o.nbytes, nn.CP.stream())
    nn.CP.synchronize()
for i in range(4):
    log.clear(); t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
    print(f'step {i}: {dt*1e3:.2f} ms, {len(log)} calls, sum host {sum(l[1] for l in log)*1e3:.2f} ms')
for name, dt, extra in sorted(log, key=lambda t: -t[1])[:12]:
    print(f'{name:28s} {dt*1e3:8.3f} ms {extra if extra else ""}')
