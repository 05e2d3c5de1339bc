"""The reference's PREDICT pipeline on the device: `PageStage` (its first stage, batched over
pages) and `TextPipeline` (the whole model system of `my_model/model.py:688-717` for one page, page -> text).

Reference (`my_model/predict.py:12-23`, `my_model/model.py:695-699`): load `model_weights.json`, pad the page to a
multiple of 16 (`make_divisible_by`), run Monochrome, feed its prediction to Paragraph, move BOTH float64 maps to the
host, where `CropAndRotateParagraphs` starts by binarising the paragraph map (`interpreter/interpreter.py:437-447`).
Here the padding, the two networks (Monochrome conv pair on tcgen05, Paragraph as one fused kernel) and the binarisation
and the connected-component labelling of that mask (`label_layer`, `interpreter/interpreter.py:16-22`) run back to
back on the device; what the crop stage needs crosses the host link as one float32 map + a uint8 mask or an int32
label map (5-8 bytes per pixel instead of 16).  `TextPipeline` continues on the device through the crop stages
(`stages.py`), Line, Char and PredToText.
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


class TextPipeline:
    """The reference's PREDICT model system (`my_model/model.py:688-717`) with every stage on the device:

        page -> pad to a multiple of 16 -> Monochrome -> Paragraph
             -> CropAndRotateParagraphs(paragraph_pred, [monochrome_pred])      (ParagraphCrop, :551-575)
             -> pad every paragraph to a multiple of 16 -> Line per paragraph    (LineSelector, :353-372)
             -> CropRotateAndZoomLines(CHAR_INPUT_HEIGHT, CHAR_FIXED_WIDTH)      (LineCrop, :595-612)
             -> Char per line                                                   (CharSelector, :375-400)
             -> PredToText                                                      (:648-657)

        pipe = TextPipeline('model_weights.json', chars=CHARS, are_similar=are_similar)
        out = pipe(page)                  # page: (1, H, W, 1) host or device array
        out['text'][paragraph_id][line_id]

    Nothing but the final hit tables (1 byte per score) and the stages' object tables (a few numbers per paragraph / line)
    crosses the host link; the reference moves every map to the host and back between the stages (`move_from_gpu_*`,
    `move_to_gpu_*`).  Networks are built per input shape and cached (the crops have data-dependent shapes), weights
    from the JSON file.  `predictors` replaces networks by callables `DeviceArray -> DeviceArray` (tests, or models kept
    elsewhere); without `chars` / `are_similar` (the reference's `primitives.CHARS`, `are_similar`) the text is returned
    as lists of class ids."""

    NETWORKS = ('monochrome', 'paragraph', 'line', 'char')

    def __init__(self, weights_path=None, find_rotation=True, predictors=None, chars=None, are_similar=None,
                 divisor=(16, 16)):
        from . import stages
        CP.use_gpu()
        self.divisor, self.chars, self.are_similar = divisor, chars, are_similar
        self._weights = weights_io.read(weights_path) if weights_path is not None else {}
        self._models, self.predictors = {}, dict(predictors or {})
        self.crop_paragraphs = stages.CropAndRotateParagraphs(None, find_rotation)
        self.crop_lines = stages.CropRotateAndZoomLines(None, my_model.CHAR_INPUT_HEIGHT, my_model.CHAR_FIXED_WIDTH)

    def predict(self, name, x):
        if name in self.predictors:
            return self.predictors[name](x)
        key = (name, tuple(x.shape))
        if key not in self._models:
            model = my_model.MAKERS[name](tuple(x.shape))
            model.set_weights(self._weights)
            self._models[key] = model
        return self._models[key].predict(x)[0]

    def to_text(self, char_pred):
        hits = glue.row_max_hits(char_pred).get()
        if self.chars is not None and self.are_similar is not None:
            return glue.hits_to_text(hits, self.chars, self.are_similar)
        rows, cols = np.nonzero(hits)
        return [int(c) for c in cols[np.argsort(rows, kind='stable')]]

    def __call__(self, page):
        x = glue.make_divisible_by(as_device(page), *self.divisor)
        monochrome_pred = self.predict('monochrome', x)
        paragraph_pred = self.predict('paragraph', monochrome_pred)
        cropped = self.crop_paragraphs(paragraph_pred, [monochrome_pred])[0]
        cropped = [glue.make_divisible_by(t, *self.divisor) for t in cropped]
        line_pred = [self.predict('line', t) for t in cropped]
        lines = self.crop_lines(line_pred, [cropped])[0]
        char_pred = [[self.predict('char', line) for line in paragraph] for paragraph in lines]
        text = [[self.to_text(pred) for pred in paragraph] for paragraph in char_pred]
        return {'text': text, 'angles': list(self.crop_paragraphs.angles), 'monochrome_pred': monochrome_pred,
                'paragraph_pred': paragraph_pred, 'cropped_monochrome': cropped, 'line_pred': line_pred,
                'cropped_2_monochrome': lines, 'char_pred': char_pred}
