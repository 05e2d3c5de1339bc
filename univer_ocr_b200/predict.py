"""First stage of the reference's PREDICT pipeline with its glue on the device.

Reference (`my_model/predict.py:12-23`, `my_model/model.py:695-699`): load `model_weights.json`, pad the page to a
multiple of 16 (`make_divisible_by`), run Monochrome, feed its prediction to Paragraph, move BOTH float64 maps to the
host, where `CropAndRotateParagraphs` starts by binarising the paragraph map (`interpreter/interpreter.py:437-447`).
Here the padding, the two networks (Monochrome conv pair on tcgen05, Paragraph as one fused kernel) and the binarisation
and the connected-component labelling of that mask (`label_layer`, `interpreter/interpreter.py:16-22`) run back to
back on the device; what the crop stage needs crosses the host link as one float32 map + a uint8 mask or an int32
label map (5-8 bytes per pixel instead of 16).  The rotation / zoom stages that follow are host code (rest of row f4).
"""
import numpy as np

from . import glue, my_model, weights_io
from .nn.gpu import CP, as_device


class PageStage:
    """
        stage = PageStage(weights_path='model_weights.json')
        out = stage(pages)                      # pages: (N, H, W, 1) host or device array, any H, W
        host = stage.to_host(out)               # {'monochrome_pred': float32, 'paragraph_mask': uint8, ...}

    Networks are (re)built per padded page shape and cached; weights come from the JSON file (or stay at their
    initialisation when the file is missing, like the reference)."""

    def __init__(self, weights_path=None, divisor=(16, 16)):
        CP.use_gpu()
        self.weights_path, self.divisor = weights_path, divisor
        self._weights = weights_io.read(weights_path) if weights_path is not None else {}
        self._models = {}

    def models_for(self, shape):
        shape = tuple(shape)
        if shape not in self._models:
            mono, para = my_model.make_monochrome(shape), my_model.make_paragraph(shape)
            for model in (mono, para):
                model.set_weights(self._weights)
            self._models[shape] = (mono, para)
        return self._models[shape]

    def __call__(self, pages):
        x = glue.make_divisible_by(as_device(pages), *self.divisor)
        mono, para = self.models_for(x.shape)
        monochrome_pred = mono.predict(x)[0]
        paragraph_pred = para.predict(monochrome_pred)[0]
        paragraph_mask = glue.thresholded(paragraph_pred)
        # what CropAndRotateParagraphs does first with that mask (interpreter/interpreter.py:437-447 -> label_layer :16-22)
        paragraph_labels, paragraph_count = glue.label_components(paragraph_mask)
        return {'padded': x, 'monochrome_pred': monochrome_pred, 'paragraph_pred': paragraph_pred,
                'paragraph_mask': paragraph_mask, 'paragraph_labels': paragraph_labels,
                'paragraph_count': paragraph_count}

    @staticmethod
    def to_host(result, want=('monochrome_pred', 'paragraph_mask')):
        return {key: np.asarray(result[key].get()) for key in want}
