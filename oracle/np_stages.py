"""CPU restatement (numpy, float64 coordinate arithmetic) of the SciPy resampling calls the reference's crop stages make
(`web_app/components/interpreter/interpreter.py`):

    ndimage.zoom(final_image, (1, zf, zf, 1), order=0)                      :514   CropRotateAndZoomLines._func2
    ndimage.rotate(array, angle, axes=(2, 1), order=0 | 1, reshape=True)    :188-192   rotate_array

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

The algorithm lives in a third-party dependency (SciPy's `ndimage`; the reference pins no version, this image has
1.18.1): `scipy/ndimage/_interpolation.py` (output shapes, rotation matrix, offsets) and `src/ni_interpolation.c`
(NI_ZoomShift, NI_GeometricTransform: coordinate arithmetic in double, nearest = floor(c + 0.5), linear = floor(c) and
weights (1 - t, t), constant mode = 0 outside [0, len - 1]).  Parity is pinned against SciPy itself:
tests/test_oracle_pin.py runs both on random inputs (bit-equal outputs required).

Arrays are (N, H, W, C); the plane that is resampled is (H, W).
"""
import numpy as np


def zoom_output_shape(shape, zoom):
    """`tuple(int(round(ii * jj)))` of ndimage.zoom (Python's round: half to even)."""
    return tuple(int(round(ii * jj)) for ii, jj in zip(shape, zoom))


def zoom_nearest(a, zoom_h, zoom_w):
    """ndimage.zoom(a, (1, zoom_h, zoom_w, 1), order=0) for a 4-D array: output index j of an axis reads input index
    floor(c + 0.5), c = j * ((in - 1) / (out - 1)) (the ratio is formed first, in double; 1 where out == 1); where the
    rounding of c leaves it above in - 1 the output is 0, the constant mode's value outside the array."""
    n, h, w, c = a.shape
    _, oh, ow, _ = zoom_output_shape(a.shape, (1, zoom_h, zoom_w, 1))

    def index(in_len, out_len):
        ratio = (in_len - 1) / (out_len - 1) if out_len > 1 else 1.0
        c = np.arange(out_len, dtype=np.float64) * ratio
        idx = np.floor(c + 0.5).astype(np.int64)
        return np.where((c < 0) | (c > in_len - 1), -1, idx)

    iy, ix = index(h, oh), index(w, ow)
    out = np.zeros((n, oh, ow, c), dtype=a.dtype)
    oky, okx = (iy >= 0) & (iy < h), (ix >= 0) & (ix < w)
    out[:, np.ix_(oky, okx)[0], np.ix_(oky, okx)[1], :] = a[:, iy[oky]][:, :, ix[okx]]
    return out


def _cosdg_sindg(angle):
    """scipy.special.cosdg / sindg: exact at multiples of 90 degrees (so quarter turns are pure index permutations)."""
    from scipy import special
    return float(special.cosdg(angle)), float(special.sindg(angle))


def rotate_geometry(h, w, angle):
    """Output plane shape, rotation matrix and offset of ndimage.rotate(..., axes=(2, 1), reshape=True) for an
    (h, w) plane: input (y, x) = matrix @ output (y, x) + offset."""
    c, s = _cosdg_sindg(angle)
    m = np.array([[c, s], [-s, c]], dtype=np.float64)
    out_bounds = m @ np.array([[0, 0, h, h], [0, w, 0, w]], dtype=np.float64)
    out_shape = (np.ptp(out_bounds, axis=1) + 0.5).astype(int)
    out_center = m @ ((out_shape - 1) / 2)
    in_center = (np.array([h, w]) - 1) / 2
    offset = in_center - out_center
    return (int(out_shape[0]), int(out_shape[1])), m, offset


def rotate(a, angle, order):
    """ndimage.rotate(a, angle, axes=(2, 1), order=order, reshape=True) for a 4-D array and order 0 (nearest) or 1
    (linear), mode 'constant', cval 0.  Coordinates: (offset + oy * m00) + ox * m01, the order in which
    NI_GeometricTransform accumulates them (found by comparing the candidates against SciPy: it decides the ties); a
    coordinate outside [0, len - 1] makes the output 0; linear weights multiply the value one axis after the other, the
    four products are summed in row-major order (double), then cast to the array's dtype."""
    assert order in (0, 1)
    n, h, w, ch = a.shape
    (oh, ow), m, off = rotate_geometry(h, w, angle)
    oy = np.arange(oh, dtype=np.float64)[:, None]
    ox = np.arange(ow, dtype=np.float64)[None, :]
    cy = (off[0] + oy * m[0, 0]) + ox * m[0, 1]
    cx = (off[1] + oy * m[1, 0]) + ox * m[1, 1]
    inside = (cy >= 0) & (cy <= h - 1) & (cx >= 0) & (cx <= w - 1)
    out = np.zeros((n, oh, ow, ch), dtype=a.dtype)
    if order == 0:
        iy = np.floor(cy + 0.5).astype(np.int64)
        ix = np.floor(cx + 0.5).astype(np.int64)
        iy, ix = np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)
        vals = a[:, iy, ix, :]
        out[:] = np.where(inside[None, :, :, None], vals, np.zeros((), dtype=a.dtype))
        return out
    fy, fx = np.floor(cy), np.floor(cx)
    ty, tx = cy - fy, cx - fx
    y0, x0 = fy.astype(np.int64), fx.astype(np.int64)

    def mirror(idx, length):                     # the neighbour past the last sample (weight 0 there): mirrored index
        if length <= 1:
            return np.zeros_like(idx)
        idx = np.where(idx < 0, -idx, idx)
        return np.where(idx >= length, 2 * length - 2 - idx, idx)

    y0c, y1c = mirror(np.clip(y0, -1, h), h), mirror(np.clip(y0 + 1, -1, h), h)
    x0c, x1c = mirror(np.clip(x0, -1, w), w), mirror(np.clip(x0 + 1, -1, w), w)
    wy = (1.0 - ty, ty)
    wx = (1.0 - tx, tx)
    t = np.zeros((n, oh, ow, ch), dtype=np.float64)
    for ia, yy in enumerate((y0c, y1c)):
        for ib, xx in enumerate((x0c, x1c)):
            coeff = a[:, yy, xx, :].astype(np.float64)
            coeff = coeff * wy[ia][None, :, :, None]
            coeff = coeff * wx[ib][None, :, :, None]
            t = t + coeff
    t = np.where(inside[None, :, :, None], t, 0.0)
    out[:] = t.astype(a.dtype)
    return out


def rotate_array(array, angle=None, good_rotation=True):
    """interpreter.py:188-192."""
    if angle is None:
        return array
    return rotate(array, angle, 1 if good_rotation else 0)


def masked_crop(image, mask, region_y, region_x):
    """`(image * mask)[:, region_y, region_x, :]` (interpreter.py:306-309)."""
    return (image * mask)[:, region_y, region_x, :]


def rotated_height(mask, angle):
    """FindObjectHeightInRotated._func (interpreter.py:229-232): rows spanned by the nearest-rotated mask."""
    rotated = rotate(mask, angle, 0)
    rows = np.flatnonzero(rotated.any(axis=(0, 2, 3)))
    return int(rows[-1] - rows[0] + 1)


def find_rotation_angle(mask, eps=1.0):
    """The ternary search of CropAndRotateSingleParagraph._func (interpreter.py:318-333)."""
    low, high = 0.0, 180.0
    while high - low > eps:
        a = low + (high - low) / 3
        b = high - (high - low) / 3
        if rotated_height(mask, a) < rotated_height(mask, b):
            high = b
        else:
            low = a
    angle = (high + low) / 2
    if not eps <= angle <= 180.0 - eps:
        angle = None
    return angle


# ---------------------------------------------------------------------------------------------- whole stages
# Restated with scipy.ndimage's label / find_objects / center_of_mass (the reference's own calls) and the resampling
# above; pinned against the reference's functions in tests/test_oracle_pin.py where /root/reference exists.
# The reference hands boolean masks to ndimage.find_objects, which the SciPy of its time read as labels 0 / 1; SciPy
# 1.18 refuses a boolean maximum label, so masks are cast to uint8 first (`_boxes`) -- same boxes.

def _boxes(mask):
    from scipy import ndimage
    return ndimage.find_objects(np.asarray(mask).astype(np.uint8))[0]


def label_layer(layer):
    """interpreter.py:16-22: one boolean mask per connected component of `layer > mean(layer)`."""
    from scipy import ndimage
    labels, count = ndimage.label(layer > np.mean(layer))
    return [labels == i + 1 for i in range(count)]


def thresholded(arr):
    """interpreter.py:437-438."""
    return arr > 0.5 * (np.mean(arr) + np.max(arr))


def crop_and_rotate_paragraphs(masks, images, find_rotation=True, eps=1.0):
    """CropAndRotateParagraphs.__call__ (:362-374) with CropAndRotateSingleParagraph._run / _func (:295-343), serially:
    -> (result[image_id][paragraph_id], angles[paragraph_id])."""
    objects = label_layer(masks)
    result = [[None] * len(objects) for _ in images]
    angles = []
    for pid, mask in enumerate(objects):
        _, ry, rx, _ = _boxes(mask)
        cropped_mask = mask[:, ry, rx, :]
        cropped = [masked_crop(image, mask, ry, rx) for image in images]
        angle = find_rotation_angle(cropped_mask, eps) if find_rotation else None
        angles.append(angle)
        rotated_mask = rotate_array(cropped_mask, angle, good_rotation=False)
        _, oy, ox, _ = _boxes(rotated_mask)
        for iid, arr in enumerate(cropped):
            result[iid][pid] = rotate_array(arr, angle)[:, oy, ox, :]
    return result, angles


def rearrange_lines(lines_top, lines_bottom):
    """interpreter.py:41-84: pairs every top mark with the nearest bottom mark (centres of mass), decides the reading
    direction from the first top / first bottom offset (scaled by 1000 until it leaves the paragraph), sorts both lists
    along it.  -> (tops, bottoms, rotation in {None, 90, 180, 270}); UnboundLocalError when no direction is decided."""
    from scipy import ndimage

    def centred(masks):
        return [(np.array(ndimage.center_of_mass(m)), m) for m in masks]

    tops, bottoms = centred(lines_top), centred(lines_bottom)
    nearest = [min(bottoms, key=lambda b: np.linalg.norm(t[0] - b[0]))[1] for t in tops]
    _, h, w, _ = lines_top[0].shape
    d = tops[0][0] - bottoms[0][0]
    while 0 < d[1] < h or 0 < d[2] < w:
        d = d * 1000
    key = rotation = None
    if abs(d[1]) > abs(d[2]):
        if d[1] < 0:
            key, rotation = (lambda item: item[0][1]), None
        elif d[1] > h:
            key, rotation = (lambda item: -item[0][1]), 180
    else:
        if d[2] < 0:
            key, rotation = (lambda item: item[0][2]), 270
        elif d[2] > w:
            key, rotation = (lambda item: -item[0][2]), 90
    if key is None:
        raise UnboundLocalError('sort_key')
    tops, bottoms = centred(lines_top), centred(nearest)
    return [t[1] for t in sorted(tops, key=key)], [b[1] for b in sorted(bottoms, key=key)], rotation


def line_region(top_mask, bottom_mask):
    """CropRotateAndZoomLines._func1 (:493-501)."""
    _, ty, tx, _ = _boxes(top_mask)
    _, by, bx, _ = _boxes(bottom_mask)
    return slice(min(ty.start, by.start), max(ty.stop, by.stop)), slice(min(tx.start, bx.start), max(tx.stop, bx.stop))


def crop_rotate_zoom(image, y, x, rotation, zoomed_height, minimal_width):
    """CropRotateAndZoomLines._func2 (:503-523)."""
    final = image[:, y, x, :]
    if rotation is not None:
        final = rotate_array(final, rotation)
    if zoomed_height is not None:
        zf = zoomed_height / final.shape[1]
        final = zoom_nearest(final, zf, zf)
    if minimal_width is not None and final.shape[2] < minimal_width:
        n, h, w, c = final.shape
        padded = np.zeros((n, h, minimal_width, c), dtype=final.dtype)
        padded[:, :, :w, :] = final
        final = padded
    return final


def crop_rotate_and_zoom_lines(masks, arrays, zoomed_height=None, minimal_width=None):
    """CropRotateAndZoomLines._func (:430-491), serially: -> result[array_id][paragraph_id][line_id]."""
    result = [[] for _ in arrays]
    for pid, mask in enumerate(masks):
        top, bottom = thresholded(mask[:, :, :, 0:1]), thresholded(mask[:, :, :, 1:2])
        tops, bottoms, rotation = rearrange_lines(label_layer(top), label_layer(bottom))
        for aid in range(len(arrays)):
            result[aid].append([])
        for t, b in zip(tops, bottoms):
            y, x = line_region(t, b)
            for aid in range(len(arrays)):
                result[aid][pid].append(crop_rotate_zoom(arrays[aid][pid], y, x, rotation, zoomed_height, minimal_width))
    return result
