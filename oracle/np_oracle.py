"""TEST INFRASTRUCTURE ONLY -- NumPy float64 restatement of the reference's layer stack.

This module is the *checker* for the CUDA path; nothing in the product package
(`univer_ocr_b200/`) imports it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may.

Every function cites the reference code it restates (paths relative to
`/root/reference/web_app/components/nn/`).  Two flavours exist where the reference uses a
Python loop over output pixels:

  * `*_loop`  -- keeps the reference's per-output-pixel loop structure (one small `dot` per
                 pixel).  This is what the reference's CPU path costs, so `bench.py` times
                 these as the `cpu_baseline` of kind "port".
  * plain     -- vectorised over pixels (one matmul per kernel tap); same arithmetic in
                 float64 (different summation order, <=1e-12 relative), used by the parity
                 tests at sizes where the loop form would take minutes.

Pinning: `tests/test_oracle_pin.py` diffs every function against the unmodified reference
(imported through `oracle/ref_loader.py`) when `/root/reference` is present, and against
the committed vectors in `tests/golden/` (generated from the reference by
`tests/golden/make_golden.py`) everywhere else.

Layouts: activations NHWC, conv weights (kh, kw, Cin, Cout), FC weights (n_in + 1, n_out)
with the bias as last row -- exactly the reference's.
"""
import math

import numpy as np

EPS_OPT = 1e-8      # optimizers.py:5
EPS_LOSS = 1e-8     # losses.py:19, :36


def _pair(v):
    if isinstance(v, (int, np.integer)):
        return int(v), int(v)
    a, b = v
    return int(a), int(b)


# --------------------------------------------------------------------------------------
# Convolutional2D  (layers/convolutional.py)
# --------------------------------------------------------------------------------------

def conv2d_out_hw(h, w, kernel_size, padding=0, stride=1):
    """layers/convolutional.py:290-301 (floor mode)."""
    kh, kw = _pair(kernel_size)
    ph, pw = _pair(padding)
    sh, sw = _pair(stride)
    return (math.floor((h + 2 * ph - (kh - 1) - 1) / sh + 1),
            math.floor((w + 2 * pw - (kw - 1) - 1) / sw + 1))


def pad_hw(X, ph, pw, value=0.0):
    """layers/convolutional.py:78-82: constant fill, value may be non-zero."""
    if ph == 0 and pw == 0:
        return X
    n, h, w, c = X.shape
    out = np.full((n, h + 2 * ph, w + 2 * pw, c), float(value), dtype=np.float64)
    out[:, ph:ph + h, pw:pw + w, :] = X
    return out


def conv2d_fwd(X, w, b, padding=0, padding_value=0.0, stride=1, bias=True):
    """Vectorised restatement of `_forward_cpu`, layers/convolutional.py:62-99."""
    X = np.asarray(X, dtype=np.float64)
    kh, kw, cin, cout = w.shape
    assert X.shape[3] == cin                                    # :65
    ph, pw = _pair(padding)
    sh, sw = _pair(stride)
    ho, wo = conv2d_out_hw(X.shape[1], X.shape[2], (kh, kw), (ph, pw), (sh, sw))
    Xp = pad_hw(X, ph, pw, padding_value)
    y = np.zeros((X.shape[0], ho, wo, cout))
    for ky in range(kh):
        for kx in range(kw):
            tap = Xp[:, ky:ky + sh * (ho - 1) + 1:sh, kx:kx + sw * (wo - 1) + 1:sw, :]
            y += tap @ w[ky, kx]
    y += float(bool(bias)) * np.asarray(b, dtype=np.float64)     # :87 bias_vec = bias * ones
    return y


def conv2d_bwd(X, w, dy, padding=0, padding_value=0.0, stride=1, bias=True):
    """Vectorised restatement of `_backward_cpu`, layers/convolutional.py:101-145.

    Returns (dX, dW, db).  `X` is the *unpadded* input; the padded copy (with
    `padding_value`) enters dW exactly as the reference's saved `_mem` does (:84, :124-128).
    """
    X = np.asarray(X, dtype=np.float64)
    dy = np.asarray(dy, dtype=np.float64)
    kh, kw, cin, cout = w.shape
    ph, pw = _pair(padding)
    sh, sw = _pair(stride)
    n, h, wid, _ = X.shape
    _, ho, wo, _ = dy.shape
    Xp = pad_hw(X, ph, pw, padding_value)
    dXp = np.zeros_like(Xp)
    dW = np.zeros_like(w, dtype=np.float64)
    for ky in range(kh):
        for kx in range(kw):
            ys = slice(ky, ky + sh * (ho - 1) + 1, sh)
            xs = slice(kx, kx + sw * (wo - 1) + 1, sw)
            dW[ky, kx] = np.tensordot(Xp[:, ys, xs, :], dy, axes=([0, 1, 2], [0, 1, 2]))
            dXp[:, ys, xs, :] += dy @ w[ky, kx].T
    db = float(bool(bias)) * dy.sum(axis=(0, 1, 2))             # :136 with bias_vec (:117)
    dX = dXp[:, ph:ph + h, pw:pw + wid, :]                      # :141-142
    return dX, dW, db


def conv2d_fwd_loop(X, w, b, padding=0, padding_value=0.0, stride=1, bias=True):
    """Per-output-pixel form, as the reference executes it (layers/convolutional.py:90-96):
    for every (y, x) a (N, kh*kw*Cin + 1) patch matrix times the (kh*kw*Cin + 1, Cout)
    weight-with-bias-row matrix."""
    X = np.asarray(X, dtype=np.float64)
    kh, kw, cin, cout = w.shape
    ph, pw = _pair(padding)
    sh, sw = _pair(stride)
    n = X.shape[0]
    ho, wo = conv2d_out_hw(X.shape[1], X.shape[2], (kh, kw), (ph, pw), (sh, sw))
    Xp = pad_hw(X, ph, pw, padding_value)
    wb = np.vstack([w.reshape(kh * kw * cin, cout), np.reshape(b, (1, cout))])
    ones = float(bool(bias)) * np.ones((n, 1))
    y = np.zeros((n, ho, wo, cout))
    for oy in range(ho):
        for ox in range(wo):
            patch = Xp[:, oy * sh:oy * sh + kh, ox * sw:ox * sw + kw, :].reshape(n, -1)
            y[:, oy, ox, :] = np.hstack([patch, ones]) @ wb
    return y


def conv2d_bwd_loop(X, w, dy, padding=0, padding_value=0.0, stride=1, bias=True):
    """Per-output-pixel form of layers/convolutional.py:121-142."""
    X = np.asarray(X, dtype=np.float64)
    kh, kw, cin, cout = w.shape
    ph, pw = _pair(padding)
    sh, sw = _pair(stride)
    n, h, wid, _ = X.shape
    _, ho, wo, _ = dy.shape
    Xp = pad_hw(X, ph, pw, padding_value)
    w2t = w.reshape(kh * kw * cin, cout).T
    ones = float(bool(bias)) * np.ones((n, 1))
    dXp = np.zeros_like(Xp)
    dwb = np.zeros((kh * kw * cin + 1, cout))
    for oy in range(ho):
        for ox in range(wo):
            g = dy[:, oy, ox, :]
            patch = Xp[:, oy * sh:oy * sh + kh, ox * sw:ox * sw + kw, :].reshape(n, -1)
            dwb += np.hstack([patch, ones]).T @ g
            dXp[:, oy * sh:oy * sh + kh, ox * sw:ox * sw + kw, :] += \
                (g @ w2t).reshape(n, kh, kw, cin)
    return (dXp[:, ph:ph + h, pw:pw + wid, :], dwb[:-1].reshape(w.shape), dwb[-1].copy())


# --------------------------------------------------------------------------------------
# Conv2DToBatchedFixedWidthed  (layers/convolutional.py:330-373)
# --------------------------------------------------------------------------------------

def window_batch_fwd(X, width):
    """Zero-pad W by `width` (left width//2), emit every width-`width` window as a batch row
    (:337-347).  (N, H, W, C) -> (N*W, H, width, C)."""
    X = np.asarray(X, dtype=np.float64)
    n, h, w, c = X.shape
    assert w >= width                                            # :365-367
    hw = width // 2
    padded = np.zeros((n, h, w + width, c))
    padded[:, :, hw:hw + w, :] = X
    win = np.lib.stride_tricks.sliding_window_view(padded, width, axis=2)   # (n,h,w+1,c,width)
    win = win[:, :, :w]                                                     # windows 0..w-1
    return np.ascontiguousarray(win.transpose(0, 2, 1, 4, 3)).reshape(n * w, h, width, c)


def window_batch_bwd(grad, in_shape, width):
    """Overlap-add of the window gradients, cropped back (:350-360)."""
    grad = np.asarray(grad, dtype=np.float64)
    n, h, w, c = in_shape
    hw = width // 2
    dxp = np.zeros((n, h, w + width, c))
    g = grad.reshape(n, w, h, width, c)
    for k in range(width):
        dxp[:, :, k:k + w, :] += g[:, :, :, k, :].transpose(0, 2, 1, 3)
    return dxp[:, :, hw:hw + w, :]


# --------------------------------------------------------------------------------------
# MaxPool2D  (layers/maxpool.py, CPU path :24-90 is the oracle)
# --------------------------------------------------------------------------------------

def maxpool2d_out_hw(h, w, kernel_size, padding=0, stride=None, ceil_mode=False):
    """layers/maxpool.py:204-216."""
    kh, kw = _pair(kernel_size)
    ph, pw = _pair(padding)
    sh, sw = (kh, kw) if stride is None else _pair(stride)
    rnd = math.ceil if ceil_mode else math.floor
    return (rnd((h + 2 * ph - (kh - 1) - 1) / sh + 1), rnd((w + 2 * pw - (kw - 1) - 1) / sw + 1))


def maxpool2d_fwd(X, kernel_size, padding=0, stride=None, ceil_mode=False):
    """layers/maxpool.py:24-57.  Zero padding (value 0 takes part in the max); windows that
    overhang the padded array (ceil_mode) are clipped.  Returns (y, mask) where mask is the
    reference's (N, kh*Ho, kw*Wo, C) tie mask (1.0 where the tap equals the window max)."""
    X = np.asarray(X, dtype=np.float64)
    kh, kw = _pair(kernel_size)
    ph, pw = _pair(padding)
    sh, sw = (kh, kw) if stride is None else _pair(stride)
    n, h, w, c = X.shape
    ho, wo = maxpool2d_out_hw(h, w, (kh, kw), (ph, pw), (sh, sw), ceil_mode)
    Xp = pad_hw(X, ph, pw, 0.0)
    y = np.zeros((n, ho, wo, c))
    mask = np.zeros((n, kh * ho, kw * wo, c))
    for oy in range(ho):
        for ox in range(wo):
            win = Xp[:, oy * sh:oy * sh + kh, ox * sw:ox * sw + kw, :]
            m = win.max(axis=(1, 2))
            sub = (win == m[:, None, None, :])
            mask[:, kh * oy:kh * oy + sub.shape[1], kw * ox:kw * ox + sub.shape[2], :] = sub
            y[:, oy, ox, :] = m
    return y, mask


def maxpool2d_bwd(grad, mask, in_shape, kernel_size, padding=0, stride=None):
    """layers/maxpool.py:61-88: each window's gradient is split equally among its tied maxima
    (ties on zero-padding taps count and their share is cropped away)."""
    grad = np.asarray(grad, dtype=np.float64)
    kh, kw = _pair(kernel_size)
    ph, pw = _pair(padding)
    sh, sw = (kh, kw) if stride is None else _pair(stride)
    n, h, w, c = in_shape
    hp, wp = h + 2 * ph, w + 2 * pw
    _, ho, wo, _ = grad.shape
    dxp = np.zeros((n, hp, wp, c))
    for oy in range(ho):
        for ox in range(wo):
            tgt = dxp[:, oy * sh:oy * sh + kh, ox * sw:ox * sw + kw, :]
            sub = mask[:, kh * oy:kh * oy + tgt.shape[1], kw * ox:kw * ox + tgt.shape[2], :]
            cnt = sub.sum(axis=(1, 2))
            tgt += (grad[:, oy, ox, :] / cnt)[:, None, None, :] * sub
    return dxp[:, ph:ph + h, pw:pw + w, :]


# --------------------------------------------------------------------------------------
# Upsample2D  (layers/upsample.py:21-39)
# --------------------------------------------------------------------------------------

def upsample2d_fwd(X, scale_factor):
    sy, sx = _pair(scale_factor)
    return np.asarray(X, dtype=np.float64).repeat(sy, axis=1).repeat(sx, axis=2)   # :24


def upsample2d_bwd(grad, scale_factor):
    """Sum over each sy x sx block (:27-38)."""
    sy, sx = _pair(scale_factor)
    grad = np.asarray(grad, dtype=np.float64)
    n, hh, ww, c = grad.shape
    return grad.reshape(n, hh // sy, sy, ww // sx, sx, c).sum(axis=(2, 4))


def upsample2d_bwd_loop(grad, scale_factor):
    sy, sx = _pair(scale_factor)
    grad = np.asarray(grad, dtype=np.float64)
    n, hh, ww, c = grad.shape
    out = np.zeros((n, hh // sy, ww // sx, c))
    for y in range(hh // sy):
        for x in range(ww // sx):
            out[:, y, x, :] += grad[:, y * sy:(y + 1) * sy, x * sx:(x + 1) * sx, :].sum(axis=(1, 2))
    return out


# --------------------------------------------------------------------------------------
# Activations / FullyConnected  (layers/layers.py)
# --------------------------------------------------------------------------------------

def leaky_relu_mask(X, alpha):
    X = np.asarray(X, dtype=np.float64)
    return (X >= 0) + alpha * (X < 0)                            # layers.py:396 (Relu: alpha=0, :379)


def leaky_relu_fwd(X, alpha=0.01):
    return X * leaky_relu_mask(X, alpha)                          # layers.py:397


def leaky_relu_bwd(X, grad, alpha=0.01):
    return grad * leaky_relu_mask(X, alpha)                       # layers.py:400


def relu_fwd(X):
    return leaky_relu_fwd(X, 0.0)


def relu_bwd(X, grad):
    return leaky_relu_bwd(X, grad, 0.0)


def sigmoid_fwd(X):
    return 1 / (1 + np.exp(-np.asarray(X, dtype=np.float64)))    # layers.py:410


def sigmoid_bwd(X, grad):
    e = np.exp(-np.asarray(X, dtype=np.float64))                  # layers.py:413-415
    return grad * e / (e + 1) ** 2


def fc_fwd(X, W):
    """y = [X, 1] . W (bias = last row of W), layers.py:335-339."""
    X = np.asarray(X, dtype=np.float64)
    return X @ W[:-1] + W[-1]


def fc_bwd(X, W, grad):
    """layers.py:341-347 -> (dX, dW)."""
    X = np.asarray(X, dtype=np.float64)
    grad = np.asarray(grad, dtype=np.float64)
    dW = np.vstack([X.T @ grad, grad.sum(axis=0, keepdims=True)])
    return grad @ W[:-1].T, dW


def concat_fwd(inputs, axis=-1):
    return np.concatenate(inputs, axis=axis)                      # layers.py:252


def concat_bwd(grad, shapes, axis=-1):
    """Slices of the gradient per input (layers.py:256-269)."""
    out, pos = [], 0
    for shp in shapes:
        idx = [slice(None)] * grad.ndim
        idx[axis] = slice(pos, pos + shp[axis])
        out.append(grad[tuple(idx)])
        pos += shp[axis]
    return out


# --------------------------------------------------------------------------------------
# Losses  (losses.py)
# --------------------------------------------------------------------------------------

def dice_loss(pred, gt):
    """SegmentationDice2D, losses.py:12-25 -> (float loss, grad)."""
    pred = np.asarray(pred, dtype=np.float64)
    gt = np.asarray(gt, dtype=np.float64)
    num = (pred * gt).sum(axis=(1, 2), keepdims=True) + EPS_LOSS
    den = pred.sum(axis=(1, 2), keepdims=True) + gt.sum(axis=(1, 2), keepdims=True) + 2 * EPS_LOSS
    loss = np.sum(1 - 2 * num / den)
    grad = -2 * (gt * den - num) / den ** 2
    return float(loss), grad


def jaccard_loss(pred, gt):
    """SegmentationJaccard2D, losses.py:29-42."""
    pred = np.asarray(pred, dtype=np.float64)
    gt = np.asarray(gt, dtype=np.float64)
    num = (pred * gt).sum(axis=(1, 2), keepdims=True) + EPS_LOSS
    den = (pred.sum(axis=(1, 2), keepdims=True) + gt.sum(axis=(1, 2), keepdims=True)
           - num + 2 * EPS_LOSS)
    loss = np.sum(1 - num / den)
    grad = -(gt * den - num * (1 - gt)) / den ** 2
    return float(loss), grad


def sigmoid_ce_loss(pred, gt):
    """SigmoidCrossEntropy, losses.py:48-57."""
    gt = np.asarray(gt, dtype=np.float64)
    p = sigmoid_fwd(pred)
    bs = gt.shape[0]
    with np.errstate(divide='ignore', invalid='ignore'):
        loss = -np.sum(gt * np.log(p) + (1 - gt) * np.log(1 - p)) / bs
    grad = (gt * (p - 1) + (1 - gt) * p) / bs
    return float(loss), grad


def softmax(X):
    X = np.asarray(X, dtype=np.float64)
    e = np.exp(X - X.max(axis=1, keepdims=True))                 # losses.py:65
    return e / e.sum(axis=1, keepdims=True)


def softmax_ce_loss(pred, gt):
    """SoftmaxCrossEntropy, losses.py:63-73.  0*log(0) = NaN is reference behaviour."""
    gt = np.asarray(gt, dtype=np.float64)
    p = softmax(pred)
    bs = gt.shape[0]
    with np.errstate(divide='ignore', invalid='ignore'):
        loss = -np.sum(gt * np.log(p)) / bs
    return float(loss), (p - gt) / bs


# --------------------------------------------------------------------------------------
# Regularisers / optimisers  (regularizations.py, optimizers.py)
# --------------------------------------------------------------------------------------

def l1_reg(w, strength):
    w = np.asarray(w, dtype=np.float64)
    return float(strength * np.sum(np.abs(w))), strength * np.sign(w)      # regularizations.py:16-19


def l2_reg(w, strength):
    w = np.asarray(w, dtype=np.float64)
    return float(strength * np.sum(w ** 2)), strength * 2 * w               # regularizations.py:23-26


def adam_update(value, grad, velocity, accumulated, lr=0.001, beta1=0.9, beta2=0.999):
    """optimizers.py:56-61.  No bias correction, eps outside the sqrt.  Returns new
    (value, velocity, accumulated)."""
    velocity = beta1 * velocity + (1 - beta1) * grad
    accumulated = beta2 * accumulated + (1 - beta2) * grad ** 2
    value = value - lr / (np.sqrt(accumulated) + EPS_OPT) * velocity
    return value, velocity, accumulated


def momentum_update(value, grad, velocity, lr, momentum=0.0):
    velocity = momentum * velocity - lr * grad                    # optimizers.py:77-79
    return value + velocity, velocity


def rmsprop_update(value, grad, accumulated, lr=0.01, rho=0.99):
    accumulated = rho * accumulated + (1 - rho) * grad ** 2      # optimizers.py:93-96
    return value - lr / (np.sqrt(accumulated) + EPS_OPT) * grad, accumulated


# --------------------------------------------------------------------------------------
# Glue with exact-index outputs  (my_model/model.py:26-34, interpreter/interpreter.py:595-614)
# --------------------------------------------------------------------------------------

def make_divisible_by(arr, y, x):
    """my_model/model.py:26-34: adds a *full* y / x when already divisible."""
    b, h, w, c = arr.shape
    ay, ax = y - h % y, x - w % x
    out = np.zeros((b, h + ay, w + ax, c))
    out[:, ay // 2:ay // 2 + h, ax // 2:ax // 2 + w, :] = arr
    return out


def pred_to_ids(pred):
    """The index part of PredToText._func1 (interpreter.py:596-602): for every row of the
    (W, n_chars) prediction, the column indices equal to the row max (ties keep *all*
    columns, in ascending order), rows whose max == 0.0 emit nothing.  Returns a flat
    int64 array of column ids in row-major order."""
    pred = np.asarray(pred)
    mx = pred.max(axis=1, keepdims=True)
    hit = (pred == mx) & (mx != 0.0)
    return np.argwhere(hit)[:, 1].astype(np.int64)


def thresholded(arr):
    """interpreter/interpreter.py:437-438 (and :549): `arr > 0.5 * (np.mean(arr) + np.max(arr))`,
    applied by the reference to one (1, H, W, 1) channel slice at a time; here to every
    (image, channel) of an NHWC tensor.  Returns a bool array."""
    arr = np.asarray(arr, dtype=np.float64)
    out = np.zeros(arr.shape, dtype=bool)
    for n in range(arr.shape[0]):
        for c in range(arr.shape[-1]):
            sl = arr[n, ..., c]
            out[n, ..., c] = sl > 0.5 * (np.mean(sl) + np.max(sl))
    return out


def pred_to_text(pred, chars, are_similar):
    """PredToText._func1 (interpreter/interpreter.py:595-614): winners of every row in row order
    (`pred_to_ids`), id 0 clears the previous character, a character "similar" to the previous
    one (primitives/__init__.py:53-54; None is similar to nothing) is skipped."""
    result, prev = '', None
    for char_id in pred_to_ids(pred):
        if char_id == 0:
            prev = None
            continue
        cur = chars[char_id]
        if are_similar(cur, prev):
            continue
        result += cur
        prev = cur
    return result
