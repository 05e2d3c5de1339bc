"""Host wall-clock of the crop stages' phases on the bench case (tests/stage_cases.py), results left on the device."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import glue, stages
from univer_ocr_b200._lib import launch_count
from tests import stage_cases

nn.CP.use_gpu()
pred, images = stage_cases.paragraph_page(0, h=496, w=736)
d_pred, d_images = nn.CP.copy(pred), [nn.CP.copy(images[0])]
lines = [stage_cases.line_paragraph(s, None, h=128, w=512, lines=3) for s in (1, 2)]
d_masks, d_arrays = [nn.CP.copy(m) for m, _ in lines], [[nn.CP.copy(a[0]) for _, a in lines]]


def timed(name, fn, reps=20):
    fn(); nn.CP.synchronize()
    n0 = launch_count(); t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    nn.CP.synchronize()
    print(f'{name:34s} {(time.perf_counter() - t0) / reps * 1e3:7.3f} ms  {(launch_count() - n0) // reps:4d} launches', flush=True)
    return out


labels, objects = timed('label_objects(paragraph map)', lambda: stages.label_objects(d_pred))
cut = timed('crop mask + map per paragraph', lambda: [(stages.crop_label_mask(labels, i + 1, *o['slices']), stages.crop(d_images[0], *o['slices'], labels, i + 1)) for i, o in enumerate(objects)])
angles = timed('angle search (2 paragraphs)', lambda: stages.find_rotation_angles([m for m, _ in cut]))
timed('rotate + box + rotate + crop', lambda: [stages.crop(stages.rotate_array(a, ang), *stages.mask_bbox(stages.rotate_array(m, ang, good_rotation=False))) for (m, a), ang in zip(cut, angles)])
timed('CropAndRotateParagraphs total', lambda: stages.CropAndRotateParagraphs(None, True)(d_pred, d_images))
marks = timed('thresholded(line map)', lambda: glue.thresholded(d_masks[0]))
timed('channel + label_objects x2', lambda: [stages.label_objects(glue.channel(marks, k)) for k in (0, 1)])
timed('CropRotateAndZoomLines total', lambda: stages.CropRotateAndZoomLines(None, 32, 8)(d_masks, d_arrays))
x = np.zeros(4, np.int32)
small = nn.DeviceArray.empty((4,), np.int32)
timed('one 16-byte read-back', lambda: small.get())
timed('one tiny launch (channel slice)', lambda: glue.channel(marks, 0))
