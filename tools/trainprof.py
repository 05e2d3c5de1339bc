"""Per-layer device time of one training step of each sub-network (CUDA event tracker)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import my_model
from univer_ocr_b200.parallel import DataParallel
from univer_ocr_b200.nn.progress_tracker import CudaEventTracker

nn.CP.use_gpu(); nn.CP.set_math_mode(sys.argv[1] if len(sys.argv) > 1 else 'tf32')
B = 64
rng = np.random.default_rng(0)
shapes = {'monochrome': (B, 496, 736, 1), 'paragraph': (B, 496, 736, 1), 'line': (B, 128, 256, 1), 'char': (B, 32, 256, 1)}
opt = nn.optimizers.Adam(lr=0.0015)
grand = 0.0
for name, shape in shapes.items():
    model = my_model.MAKERS[name](shape, optimizer=opt)
    dp = DataParallel(model, optimizer=opt)
    X = nn.CP.copy(rng.random(shape, dtype=np.float32))
    out_shape = model.get_output_shapes([shape])[0]
    if name == 'char':
        y = np.zeros(out_shape, dtype=np.float32); y[np.arange(out_shape[0]), rng.integers(0, out_shape[1], out_shape[0])] = 1
    else:
        y = (rng.random(out_shape, dtype=np.float32) < 0.2).astype(np.float32)
    y = nn.CP.copy(y)
    for _ in range(2): dp.train(X, y)
    tracker = CudaEventTracker()
    for layer in model.layers.values(): layer.progress_tracker = tracker
    import ctypes
    from univer_ocr_b200._lib import lib
    def ev():
        e = ctypes.c_void_p(); lib.uocr_event_create(ctypes.byref(e)); return e.value
    e0, e1 = ev(), ev()
    lib.uocr_event_record(e0, nn.CP.stream())
    dp.train(X, y)
    lib.uocr_event_record(e1, nn.CP.stream()); lib.uocr_event_sync(e1)
    ms = ctypes.c_float(0); lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    rows = tracker.summary_ms()
    tot = sum(v[0] for v in rows.values())
    grand += ms.value
    print(f'== {name}: step {ms.value:.3f} ms, tracked layers {tot:.3f} ms')
    for (lname, evn), v in sorted(rows.items(), key=lambda kv: -kv[1][0])[:10]:
        print(f'   {v[0]:8.3f} ms  {evn:8s} {lname}')
print('total', grand)
