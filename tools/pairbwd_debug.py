import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200._lib import ACT_LEAKY, ACT_NONE, lib
nn.CP.use_gpu()
rng = np.random.default_rng(29)
for (n, h, w), act1 in (((1, 4, 8), ACT_NONE), ((1, 4, 8), ACT_LEAKY), ((1, 16, 256), ACT_LEAKY), ((2, 16, 256), ACT_LEAKY)):
    X = rng.uniform(size=(n, h, w, 1)).astype(np.float32)
    w1 = (rng.standard_normal((3, 3, 1, 16)) * 0.4).astype(np.float32)
    b1 = (rng.standard_normal(16) * 0.2).astype(np.float32)
    w2 = (rng.standard_normal((3, 3, 16, 1)) * 0.3).astype(np.float32)
    dy = rng.standard_normal((n, h, w, 1)).astype(np.float32)
    d = [nn.CP.copy(a) for a in (X, w1, b1, w2, dy)]
    need = ctypes.c_size_t(0)
    lib.uocr_conv3x3_pair_bwd_workspace(n, h, w, 16, ctypes.byref(need))
    outs = []
    for mode in (0, 1):
        ws = nn.DeviceArray(((need.value + 3) // 4,))
        g = [nn.DeviceArray.zeros(s_) for s_ in ((3, 3, 1, 16), (16,), (3, 3, 16, 1), (1,))]
        lib.uocr_conv3x3_pair_bwd_mode(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, None, g[0].ptr, g[1].ptr,
                                       g[2].ptr, g[3].ptr, n, h, w, 16, act1, 0.01, 0, ws.ptr, need.value, mode, nn.CP.stream())
        outs.append([np.asarray(t.get(), dtype=np.float64) for t in g])
    print('case', (n, h, w), 'act', act1)
    for name, a, b in zip(('dw1', 'db1', 'dw2', 'db2'), outs[1], outs[0]):
        err = np.abs(a - b)
        print(f'  {name}: max err {err.max():.3e} / max {np.abs(b).max():.3e}; worst idx {np.unravel_index(err.argmax(), err.shape)}')
    if h * w <= 64:
        print('  dw1 tc  ch0', outs[1][0][:, :, 0, 0].round(4).tolist())
        print('  dw1 ref ch0', outs[0][0][:, :, 0, 0].round(4).tolist())
