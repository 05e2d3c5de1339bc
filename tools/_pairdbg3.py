import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200._lib import ACT_LEAKY, ACT_SIGMOID, lib
nn.CP.use_gpu()
mode = sys.argv[1]
rng = np.random.default_rng(5)
n, h, w = 64, 496, 736
X = rng.uniform(size=(n, h, w, 1)).astype(np.float32)
w1 = (rng.standard_normal((3, 3, 1, 16)) * 0.4).astype(np.float32); b1 = (rng.standard_normal(16) * 0.2).astype(np.float32)
w2 = (rng.standard_normal((3, 3, 16, 1)) * 0.3).astype(np.float32); b2 = rng.standard_normal(1).astype(np.float32)
d = [nn.CP.copy(a) for a in (X, w1, b1, w2, b2)]
st = nn.CP.stream()
flush = nn.DeviceArray((64 * 1024 * 1024,))
def launch(y):
    lib.uocr_conv3x3_pair_fwd(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, y.ptr, n, h, w, 16, ACT_LEAKY, 0.01, ACT_SIGMOID, 0.0, 1, st)
try:
    seq = {'a': ['2', '3'], 'b': ['3'], 'c': ['3', '2', '3']}[mode[0]]
    for tc in seq:
        os.environ['UOCR_PAIR_TC'] = tc
        y = nn.DeviceArray((n, h, w, 1))
        for i in range(3): launch(y)
        for i in range(5):
            if 'f' in mode: flush.fill(0)
            launch(y)
            nn.CP.synchronize()
        print('tc', tc, 'ok mean', float(y.get().mean()), flush=True)
except Exception as e:
    print('FAILED', mode, str(e)[:100])
