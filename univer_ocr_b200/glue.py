"""Device-side glue between the sub-networks.

In the reference's PREDICT pipeline every prediction goes back to the host as float64 before the
next stage looks at it (`my_model/model.py:695-714`: `move_from_gpu_*` components), although the
stages only need (a) the page padded to a multiple of 16, (b) a binarised mask of each predicted
map, (c) which characters won each window.  These three run on the device here, so what crosses
the host link is uint8 masks / hit tables (1 byte per element instead of 8):

    pixels_to_unit      my_model/train_data_generator.py:24-37  (uint8 planes / 255, widened on the device)
    make_divisible_by   my_model/model.py:26-34
    thresholded         interpreter/interpreter.py:437-438, 549   (arr > 0.5 * (mean + max))
    row_max_hits / pred_to_text   interpreter/interpreter.py:595-614  (PredToText._func1)
    label_components / label_layer   interpreter/interpreter.py:16-22  (ndimage.label of the crop stages)

The stages between them (row f4: label, crop, rotate, zoom) are in `stages.py`, built on the labelling and statistics
here.
"""
import ctypes

import numpy as np

from ._lib import lib
from .nn.gpu import DeviceArray, as_device, stream


def make_divisible_by(arr, y, x):
    """Zero-pads H and W of an NHWC device tensor up to the next multiple of (y, x); a dimension
    that is already aligned still grows by a full y / x, and the extra rows are split top = add // 2
    like the reference's `np.pad(((0, 0), (add_y // 2, add_y - add_y // 2), ...))`."""
    arr = as_device(arr)
    n, h, w, c = arr.shape
    add_y, add_x = y - h % y, x - w % x
    top, left = add_y // 2, add_x // 2
    out = DeviceArray.empty((n, h + add_y, w + add_x, c))
    lib.uocr_pad_hw_f32(out.ptr, arr.ptr, n, h, w, c, top, add_y - top, left, add_x - left, 0.0, stream())
    return out


def pixels_to_unit(arr, divisor=255.0):
    """uint8 image planes (device or host, any shape) -> float32 `pixel / 255`, bit-identical to the float32 storage of
    the reference's `encode_layers` output (`train_data_generator.py:24-37`: PNG planes / 255 in float64).  Uploading
    the bytes and widening them here moves a quarter of the float32 payload over the host link."""
    arr = as_device(arr) if isinstance(arr, DeviceArray) else DeviceArray.from_host(np.asarray(arr, np.uint8), np.uint8)
    assert arr.dtype == np.uint8, f'expected uint8 pixels, got {arr.dtype}'
    out = DeviceArray.empty(arr.shape, np.float32)
    lib.uocr_u8_div_f32(out.ptr, arr.ptr, float(divisor), arr.size, stream())
    return out


def thresholded(arr):
    """uint8 mask of `arr > 0.5 * (mean + max)`, mean and max taken per (image, channel) over H x W
    -- the reference thresholds one (1, H, W, 1) channel slice at a time."""
    arr = as_device(arr)
    n, c = arr.shape[0], arr.shape[-1]
    hw = arr.size // (n * c)
    mask = DeviceArray.empty(arr.shape, np.uint8)
    nbytes = ctypes.c_size_t(0)
    lib.uocr_threshold_mask_workspace(n, c, ctypes.byref(nbytes))
    work = DeviceArray.empty((nbytes.value,), np.uint8)
    lib.uocr_threshold_mask(arr.ptr, mask.ptr, n, hw, c, work.ptr, stream())
    return mask


def above_mean(arr):
    """uint8 mask of `arr > mean(arr)`, the mean taken per (image, channel) over H x W in float64: the foreground
    `label_layer` labels when it is handed a float map (`interpreter/interpreter.py:16-17`)."""
    arr = as_device(arr)
    n, c = arr.shape[0], arr.shape[-1]
    hw = arr.size // (n * c)
    mask = DeviceArray.empty(arr.shape, np.uint8)
    nbytes = ctypes.c_size_t(0)
    lib.uocr_threshold_mask_workspace(n, c, ctypes.byref(nbytes))
    work = DeviceArray.empty((nbytes.value,), np.uint8)
    lib.uocr_above_mean_mask(arr.ptr, mask.ptr, n, hw, c, work.ptr, stream())
    return mask


def channel(mask, k):
    """`mask[:, :, :, k:k+1]` of a uint8 (N, H, W, C) device mask, contiguous."""
    n, h, w, c = mask.shape
    out = DeviceArray.empty((n, h, w, 1), np.uint8)
    lib.uocr_channel_slice_u8(mask.ptr, out.ptr, n * h * w, c, k, stream())
    return out


def channel_planes(mask):
    """(1, H, W, C) uint8 device mask -> (C, H, W, 1): the channels as separate images (`mask[:, :, :, k:k+1]` for every
    k at once), e.g. to label the top and bottom marks of a Line mask in one `label_components` call."""
    n, h, w, c = mask.shape
    assert n == 1, f'expected one image, got {n}'
    out = DeviceArray.empty((c, h, w, 1), np.uint8)
    lib.uocr_channel_planes_u8(mask.ptr, out.ptr, h * w, c, stream())
    return out


def label_components(mask):
    """Connected components of every image of a uint8 mask (N, H, W) or (N, H, W, 1) -> (labels int32 of the same
    shape, counts int32 (N,)), both on the device: foreground = strictly above the image's mean, 4-neighbourhood,
    labels 1..count in raster order of each component's first pixel -- `ndimage.label(layer > np.mean(layer))` of the
    reference's `label_layer` (`interpreter/interpreter.py:16-22`), bit for bit."""
    mask = as_device(mask) if isinstance(mask, DeviceArray) else DeviceArray.from_host(np.asarray(mask, np.uint8), np.uint8)
    assert mask.dtype == np.uint8, f'expected a uint8 mask, got {mask.dtype}'
    shape = mask.shape
    assert len(shape) == 3 or (len(shape) == 4 and shape[3] == 1), f'expected (N, H, W[, 1]), got {shape}'
    n, h, w = shape[:3]
    nbytes = ctypes.c_size_t(0)
    lib.uocr_label_components_workspace(n, h, w, ctypes.byref(nbytes))
    work = DeviceArray.empty((nbytes.value,), np.uint8)
    labels = DeviceArray.empty(shape, np.int32)
    counts = DeviceArray.empty((n,), np.int32)
    lib.uocr_label_components(mask.ptr, labels.ptr, counts.ptr, n, h, w, work.ptr, stream())
    return labels, counts


def label_stats(labels, counts):
    """Per-object bounding boxes and centres of mass of a device label map (N, H, W[, 1]) from `label_components`:
    -> list (per image) of lists (per object, in label order) of dicts {'slices': (slice_y, slice_x) as
    `ndimage.find_objects` returns them, 'center_of_mass': (y, x) as `ndimage.center_of_mass` of the object's mask,
    'count': pixels}.  The sums are accumulated as integers on the device; one small table crosses the host link."""
    n, h, w = labels.shape[:3]
    host_counts = np.asarray(counts.get(), dtype=np.int64)
    max_labels = max(int(host_counts.max()) if host_counts.size else 0, 1)
    stats = DeviceArray.empty((n, max_labels, 7), np.int64)
    lib.uocr_label_stats(labels.ptr, stats.ptr, n, h, w, max_labels, stream())
    table = stats.get()
    out = []
    for i in range(n):
        objs = []
        for l in range(int(host_counts[i])):
            cnt, sy, sx, y0, y1, x0, x1 = (int(v) for v in table[i, l])
            objs.append({'slices': (slice(y0, y1 + 1), slice(x0, x1 + 1)),
                         'center_of_mass': (sy / cnt, sx / cnt), 'count': cnt})
        out.append(objs)
    return out


def label_layer(mask):
    """The reference's `label_layer` (`interpreter/interpreter.py:16-22`) for ONE (1, H, W, 1) mask: a list of
    full-size boolean arrays, one per object.  The labelling runs on the device; the per-object masks are expanded on
    the host (that list is the reference's interface -- device-side consumers use `label_components`)."""
    labels, counts = label_components(mask)
    host = labels.get()
    return [host == l_id + 1 for l_id in range(int(counts.get()[0]))]


def row_max_hits(pred):
    """uint8 (rows, classes): 1 where the class ties the row maximum and that maximum is not 0."""
    pred = as_device(pred)
    rows, cols = pred.shape
    hits = DeviceArray.empty((rows, cols), np.uint8)
    lib.uocr_row_max_hits(pred.ptr, hits.ptr, rows, cols, stream())
    return hits


def hits_to_text(hits, chars, are_similar):
    """The string part of PredToText._func1 (:602-614) on a host hit table: class ids in row
    order (ties: every hit of the row, ascending), id 0 resets the repeat filter, a character
    similar to the previous one is dropped."""
    rows, cols = np.nonzero(np.asarray(hits))
    result, prev = '', None
    for char_id in cols[np.argsort(rows, kind='stable')]:
        if char_id == 0:
            prev = None
            continue
        cur = chars[char_id]
        if are_similar(cur, prev):
            continue
        result += cur
        prev = cur
    return result


def pred_to_text(pred, chars, are_similar):
    """Char-network scores (rows, classes) on the device → text: hit table on the device, 1 byte
    per score over the host link, string assembly on the host."""
    return hits_to_text(row_max_hits(pred).get(), chars, are_similar)
