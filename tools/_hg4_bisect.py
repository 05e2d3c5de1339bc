import os, subprocess, sys
code = r'''
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from univer_ocr_b200 import my_model
from univer_ocr_b200.nn.gpu import CP
import bench
CP.use_gpu(); CP.set_math_mode('tf32')
shape = (2, 64, 256, 1)
m = my_model.make_line(shape); bench.signed_init(m, 3.5, True)
x = np.random.default_rng(1).uniform(0, 1, shape).astype(np.float32)
y = m.predict(x)[0].get()
print('ok', float(np.abs(y).max()))
'''
for stop, dbg in (('0', '0'), ('1', '0'), ('2', '0'), ('3', '0'), ('4', '0'), ('-1', '0')):
    env = dict(os.environ, UOCR_HG4_STOP=stop, UOCR_HG4_DBG=dbg)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=120)
    print('stop', stop, 'dbg', dbg, (r.stdout.strip().splitlines() or ['-'])[-1], '|', (r.stderr.strip().splitlines() or ['-'])[-1][-150:], flush=True)
