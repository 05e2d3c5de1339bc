"""Four-channel hourglass kernel (uocr_hourglass4_fwd) against the layer-by-layer paths of the same Line network.

    python tools/hg4_check.py            # parity at several geometries + timing at batch 64
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from univer_ocr_b200 import my_model                      # noqa: E402
from univer_ocr_b200.nn.gpu import CP                     # noqa: E402
set_math_mode = CP.set_math_mode


def signed(model, seed, scale=3.5):
    """bench.signed_init: centred weights and biases, so that the sigmoids are not saturated."""
    import bench
    bench.signed_init(model, scale, True)


def main():
    CP.use_gpu()
    worst = 0.0
    for shape in ((1, 16, 128, 1), (2, 32, 256, 1), (1, 20, 36, 1), (3, 128, 256, 1), (1, 52, 300, 1), (2, 4, 4, 1)):
        model = my_model.make_line(shape)
        signed(model, 7)
        x = np.random.default_rng(1).uniform(0, 1, shape).astype(np.float32)
        set_math_mode('tf32')
        fusion = model.infer_fusion
        assert fusion is not None, 'the Line network was not recognised'
        y_fused = model.predict(x)[0].get()
        model.infer_fusion = None
        y_tf32 = model.predict(x)[0].get()
        set_math_mode('fp32')
        y_fp32 = model.predict(x)[0].get()
        set_math_mode('tf32')
        e1, e2 = np.abs(y_fused - y_fp32).max(), np.abs(y_tf32 - y_fp32).max()
        worst = max(worst, e1)
        print(f'{shape}: fused vs fp32 {e1:.2e}   layerwise tf32 vs fp32 {e2:.2e}   mean {y_fp32.mean():.3f} '
              f'finite {np.isfinite(y_fused).all()}', flush=True)
    shape = (64, 128, 256, 1)
    model = my_model.make_line(shape)
    signed(model, 7)
    from univer_ocr_b200.nn.gpu import DeviceArray
    x = DeviceArray.from_host(np.random.default_rng(1).uniform(0, 1, shape).astype(np.float32)) if hasattr(DeviceArray, 'from_host') else None
    xin = x if x is not None else np.random.default_rng(1).uniform(0, 1, shape).astype(np.float32)
    for name, fusion in (('fused', model.infer_fusion), ('layerwise', None)):
        model.infer_fusion = fusion
        for _ in range(3):
            model.predict(xin)
        CP.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            model.predict(xin)
        CP.synchronize()
        print(f'{name}: {(time.perf_counter() - t0) / 50 * 1e3:.4f} ms per forward at batch 64', flush=True)
    print('worst', worst)
    return 0 if worst < 4e-3 else 1


if __name__ == '__main__':
    sys.exit(main())
