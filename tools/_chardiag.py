import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle import np_models
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import my_model
nn.CP.use_gpu()
name = sys.argv[1] if len(sys.argv) > 1 else 'char'
shape = {'char': (2, 32, 64, 1), 'line': (2, 64, 128, 1)}[name]
rng = np.random.default_rng(321)
spec, kind = np_models.net_spec(name), np_models.loss_kind(name)
w = np_models.golden_weights(name, 321)
X = rng.uniform(size=shape).astype(np.float32).astype(np.float64)
pred = np_models.forward(spec, w, X)
if kind == 'dice':
    y = (rng.uniform(size=pred.shape) < 0.2).astype(np.float64)
else:
    y = np.zeros(pred.shape); y[np.arange(y.shape[0]), rng.integers(0, y.shape[1], size=y.shape[0])] = 1
rec = {}
for mode in ('fp32', 'tf32'):
    nn.CP.set_math_mode(mode)
    model = my_model.MAKERS[name](shape, optimizer=nn.optimizers.Adam(lr=0.0015))
    model.set_weights({k: {n: v.tolist() for n, v in p.items()} for k, p in w.items()})
    r = rec[mode] = {}
    for lname, layer in model.layers.items():
        def wrap(layer=layer, lname=lname, orig=layer.backward):
            def bw(grads):
                g = grads[0] if isinstance(grads, list) else grads
                r[lname + ' IN'] = np.asarray(nn.gpu.as_device(g).get(), dtype=np.float64)
                out = orig(grads)
                o = out[0] if isinstance(out, list) else out
                if o is not None:
                    r[lname + ' OUT'] = np.asarray(o.get(), dtype=np.float64)
                return out
            return bw
        layer.backward = wrap()
    predicted = model.forward([X])
    loss, g = model._loss_for(0)(predicted[0], y)
    model.backward([g])
    for key, p in model.params().items():
        r[key + ' GRAD'] = np.asarray(p.grad.get(), dtype=np.float64)
for k in rec['fp32']:
    a, b = rec['fp32'][k], rec['tf32'].get(k)
    if b is None: print(k, 'missing in tf32'); continue
    print(f'{k:55s} max|fp32| {np.max(np.abs(a)):.3e}  rel diff {np.max(np.abs(a-b))/max(np.max(np.abs(a)),1e-30):.2e}')
