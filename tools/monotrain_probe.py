import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import my_model
from univer_ocr_b200.parallel import DataParallel
nn.CP.use_gpu(); nn.CP.set_math_mode('tf32')
B = 64
rng = np.random.default_rng(0)
shape = (B, 496, 736, 1)
opt = nn.optimizers.Adam(lr=0.0015)
model = my_model.make_monochrome(shape, optimizer=opt)
dp = DataParallel(model, optimizer=opt)
X = nn.CP.copy(rng.random(shape, dtype=np.float32))
y = nn.CP.copy((rng.random(shape, dtype=np.float32) < 0.2).astype(np.float32))
for _ in range(3): dp.train(X, y)
nn.CP.synchronize() if hasattr(nn.CP, 'synchronize') else None
print('ok')
