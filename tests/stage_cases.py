"""Synthetic inputs for the crop-stage tests (shared by the oracle pin tests and the GPU parity tests)."""
import numpy as np


def paragraph_page(seed=0, h=160, w=240, tilt=(12.0, -25.0)):
    """A Paragraph prediction (1, h, w, 1) float32 with two tilted rectangular paragraphs (values near 1 inside, near 0
    outside, noisy) and two (1, h, w, C) float32 maps to cut."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    pred = np.zeros((h, w))
    centres = ((h * 0.3, w * 0.3), (h * 0.68, w * 0.65))
    sizes = ((h * 0.16, w * 0.2), (h * 0.12, w * 0.24))
    for (cy, cx), (hy, hx), deg in zip(centres, sizes, tilt):
        a = np.deg2rad(deg)
        u = (yy - cy) * np.cos(a) + (xx - cx) * np.sin(a)
        v = -(yy - cy) * np.sin(a) + (xx - cx) * np.cos(a)
        pred[(np.abs(u) < hy) & (np.abs(v) < hx)] = 1.0
    pred = np.clip(pred * 0.9 + rng.uniform(0, 0.08, size=pred.shape), 0, 1).astype(np.float32)[None, :, :, None]
    images = [rng.uniform(0.05, 1, size=(1, h, w, 1)).astype(np.float32),
              rng.uniform(-1, 1, size=(1, h, w, 2)).astype(np.float32)]
    return pred, images


def line_paragraph(seed=0, direction=None, h=96, w=208, lines=3):
    """A Line prediction (1, h, w, 2) float32 of one paragraph with `lines` text lines (channel 0: a mark along the top of
    every line, channel 1: along its bottom) and two maps to cut.  `direction`: None (upright), 90, 180, 270 = the
    quarter turn rearrange_lines should report (the page content is turned accordingly)."""
    rng = np.random.default_rng(seed)
    top, bottom = np.zeros((h, w)), np.zeros((h, w))
    pitch = h // (lines + 1)
    for i in range(lines):
        y = pitch * (i + 1) - pitch // 4
        x0, x1 = 10 + 3 * i, w - 14 - 5 * i
        top[y - 7:y - 5, x0:x1] = 1.0
        bottom[y + 6:y + 8, x0:x1] = 1.0
    mask = np.stack([top, bottom], axis=-1)
    content = rng.uniform(0.05, 1, size=(h, w, 3))
    turns = {None: 0, 270: 1, 180: 2, 90: 3}[direction]      # np.rot90 counter-clockwise quarter turns of the page
    mask, content = np.rot90(mask, turns, axes=(0, 1)), np.rot90(content, turns, axes=(0, 1))
    mask = np.clip(mask * 0.85 + rng.uniform(0, 0.1, size=mask.shape), 0, 1)
    mask = np.ascontiguousarray(mask, dtype=np.float32)[None]
    content = np.ascontiguousarray(content, dtype=np.float32)[None]
    return mask, [content[..., 0:1].copy(), content[..., 1:3].copy()]
