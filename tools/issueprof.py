"""Is the training / inference step bound by the host issuing launches?  Compares the wall time the Python side needs
to ISSUE n steps (no sync) with the device time of the same steps."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import my_model
from univer_ocr_b200.parallel import DataParallel
from univer_ocr_b200.pipeline import ConcurrentBranches

nn.CP.use_gpu(); nn.CP.set_math_mode('tf32')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
shapes = {'monochrome': (B, 496, 736, 1), 'paragraph': (B, 496, 736, 1), 'line': (B, 128, 256, 1), 'char': (B, 32, 256, 1)}
opt = nn.optimizers.Adam(lr=0.0015)
dps, feeds, targets = {}, {}, {}
for name, shape in shapes.items():
    model = my_model.MAKERS[name](shape, optimizer=opt)
    dps[name] = DataParallel(model, optimizer=opt)
    feeds[name] = nn.CP.copy(rng.random(shape, dtype=np.float32))
    out_shape = model.get_output_shapes([shape])[0]
    if name == 'char':
        y = np.zeros(out_shape, dtype=np.float32); y[np.arange(out_shape[0]), rng.integers(0, out_shape[1], out_shape[0])] = 1
    else:
        y = (rng.random(out_shape, dtype=np.float32) < 0.2).astype(np.float32)
    targets[name] = nn.CP.copy(y)
fork = ConcurrentBranches(4)
names = list(dps)


def step_fork():
    fork.run(*[(lambda n=n: dps[n].train(feeds[n], targets[n])) for n in names])


def step_serial():
    for n in names:
        dps[n].train(feeds[n], targets[n])


for label, fn in (('forked', step_fork), ('serial', step_serial)):
    for _ in range(3):
        fn()
    nn.CP.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    nn.CP.synchronize()
    t2 = time.perf_counter()
    print(f'{label}: issue {1e3 * (t1 - t0) / n:.3f} ms/step, total {1e3 * (t2 - t0) / n:.3f} ms/step', flush=True)
for name in names:
    fn = lambda: dps[name].train(feeds[name], targets[name])
    for _ in range(3):
        fn()
    nn.CP.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    nn.CP.synchronize()
    t2 = time.perf_counter()
    print(f'{name}: issue {1e3 * (t1 - t0) / n:.3f} ms/step, total {1e3 * (t2 - t0) / n:.3f} ms/step', flush=True)
