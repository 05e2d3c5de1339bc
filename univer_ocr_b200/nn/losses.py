"""Losses (reference: nn/losses.py).  `loss(prediction, ground_truth) -> (loss, grad)` as in
the reference; `loss` is a LazyScalar (behaves like the reference's float, but the device is
only synchronised when the number is actually read)."""
import ctypes

from .._lib import SEG_DICE, SEG_JACCARD, lib
from .gpu import DeviceArray, LazyScalar, as_device, stream


class BaseLoss:
    def __call__(self, prediction, ground_truth, want_grad=True):
        raise NotImplementedError()


class _Segmentation(BaseLoss):
    kind = None

    def __call__(self, prediction, ground_truth, want_grad=True):
        pred, gt = as_device(prediction), as_device(ground_truth)
        assert pred.shape == gt.shape, f'{pred.shape} != {gt.shape}'
        n, h, w, c = pred.shape
        need = ctypes.c_size_t(0)
        lib.uocr_seg_loss_workspace(n, c, ctypes.byref(need))
        ws = DeviceArray((need.value + 7) // 8, 'float64')
        loss = DeviceArray((1,))
        grad = DeviceArray(pred.shape) if want_grad else None
        lib.uocr_seg_loss(self.kind, pred.ptr, gt.ptr, grad.ptr if want_grad else None, loss.ptr,
                          n, h * w, c, ws.ptr, stream())
        return LazyScalar(loss), grad


class SegmentationDice2D(_Segmentation):
    """losses.py:9-25"""
    kind = SEG_DICE


class SegmentationJaccard2D(_Segmentation):
    """losses.py:28-42"""
    kind = SEG_JACCARD


class SigmoidCrossEntropy(BaseLoss):
    """losses.py:45-57"""

    def __call__(self, prediction, ground_truth, want_grad=True):
        pred, gt = as_device(prediction), as_device(ground_truth)
        assert pred.shape == gt.shape, f'{pred.shape} != {gt.shape}'
        batch = gt.shape[0]
        loss = DeviceArray((1,))
        grad = DeviceArray(pred.shape) if want_grad else None
        lib.uocr_sigmoid_ce(pred.ptr, gt.ptr, grad.ptr if want_grad else None, loss.ptr, batch,
                            pred.size // batch, stream())
        return LazyScalar(loss), grad


class SoftmaxCrossEntropy(BaseLoss):
    """losses.py:60-73"""

    def __call__(self, prediction, ground_truth, want_grad=True):
        pred, gt = as_device(prediction), as_device(ground_truth)
        assert pred.shape == gt.shape and pred.ndim == 2, f'{pred.shape} vs {gt.shape}'
        batch, classes = pred.shape
        loss = DeviceArray((1,))
        ws = DeviceArray((batch,))
        grad = DeviceArray(pred.shape) if want_grad else None
        lib.uocr_softmax_ce(pred.ptr, gt.ptr, grad.ptr if want_grad else None, loss.ptr, batch,
                            classes, ws.ptr, stream())
        return LazyScalar(loss), grad
