"""Runs the Monochrome pair forward (UOCR_PAIR_TC from the environment, default 3) a few times at batch 64: the
command ncu captures (tools/pairrows_check.py does the correctness sweep)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('UOCR_PAIR_TC', '3')
import univer_ocr_b200.nn as nn
from univer_ocr_b200._lib import ACT_LEAKY, ACT_SIGMOID, lib
nn.CP.use_gpu()
rng = np.random.default_rng(5)
n, h, w = int(os.environ.get('PAIR_N', 64)), 496, 736
X = rng.uniform(size=(n, h, w, 1)).astype(np.float32)
w1 = (rng.standard_normal((3, 3, 1, 16)) * 0.4).astype(np.float32)
b1 = (rng.standard_normal(16) * 0.2).astype(np.float32)
w2 = (rng.standard_normal((3, 3, 16, 1)) * 0.3).astype(np.float32)
b2 = rng.standard_normal(1).astype(np.float32)
d = [nn.CP.copy(a) for a in (X, w1, b1, w2, b2)]
y = nn.DeviceArray((n, h, w, 1))
for _ in range(4):
    lib.uocr_conv3x3_pair_fwd(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, y.ptr, n, h, w, 16,
                              ACT_LEAKY, 0.01, ACT_SIGMOID, 0.0, 1, nn.CP.stream())
nn.CP.synchronize()
print('done', float(y.get().mean()))
