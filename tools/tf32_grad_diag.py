"""Per-parameter gradient error of one training step: TF32 mode and FP32 mode vs the float64 oracle, same weights.
Diagnostic for tests/test_gpu_round2.py::test_data_parallel_train_step_vs_oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_models
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import my_model

nn.CP.use_gpu()
SHAPES = {'monochrome': (2, 64, 96, 1), 'paragraph': (2, 64, 96, 1), 'line': (2, 64, 128, 1), 'char': (2, 32, 64, 1)}
names = sys.argv[1:] or list(SHAPES)
for name in names:
    rng = np.random.default_rng(321)
    spec, kind = np_models.net_spec(name), np_models.loss_kind(name)
    w = np_models.golden_weights(name, 321)
    X = rng.uniform(size=SHAPES[name]).astype(np.float32).astype(np.float64)
    pred, saved = np_models.forward(spec, w, X, keep=True)
    if kind == 'dice':
        y = (rng.uniform(size=pred.shape) < 0.2).astype(np.float64)
    else:
        y = np.zeros(pred.shape); y[np.arange(y.shape[0]), rng.integers(0, y.shape[1], size=y.shape[0])] = 1
    _, grad = np_models.loss_and_grad(kind, pred, y)
    _, want = np_models.backward(spec, w, saved, grad)
    for fusion in (True, False):
        for mode in ('fp32', 'tf32'):
            nn.CP.set_math_mode(mode)
            model = my_model.MAKERS[name](SHAPES[name], optimizer=nn.optimizers.Adam(lr=0.0015))
            if not fusion:
                model.fusion = False
                model.initialize(model.input_shapes)
            model.set_weights({k: {n: v.tolist() for n, v in p.items()} for k, p in w.items()})
            model.fused_update = False
            predicted = model.forward([X])
            perr = np.max(np.abs(predicted[0].get() - pred)) / np.max(np.abs(pred))
            loss, g = model._loss_for(0)(predicted[0], y)
            model.backward([g])
            out = []
            for key, p in model.params().items():
                lkey, pname = key.rsplit('/', 1)
                ww = want[lkey][pname]
                err = np.max(np.abs(p.grad.get().astype(np.float64) - ww)) / np.max(np.abs(ww))
                out.append(f'{lkey.split("/")[-1]}.{pname}={err:.1e}')
            print(f'{name:10s} fusion={int(fusion)} {mode}: pred {perr:.1e} | ' + ' '.join(out), flush=True)
