import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from univer_ocr_b200 import my_model
from univer_ocr_b200.nn.gpu import CP, DeviceArray
import bench
CP.use_gpu(); CP.set_math_mode('tf32')
shape = (64, 128, 256, 1)
m = my_model.make_line(shape); bench.signed_init(m, 3.5, True)
x = DeviceArray.from_host(np.random.default_rng(1).uniform(0, 1, shape).astype(np.float32))
for _ in range(5):
    m.predict(x)
CP.synchronize()
