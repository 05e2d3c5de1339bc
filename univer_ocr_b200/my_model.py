"""The four sub-networks of the OCR model, built from the B200 layer stack with the same
builder functions, layer names and hyper-parameters as the reference's
`web_app/components/my_model/model.py` (:37-304), so that `model_weights.json` files are
interchangeable (keys 'Monochrome/conv_1', 'Paragraph/up_2/conv_block/conv_1',
'Char/dense_block/dense_1', ...).

Only the network definitions live here -- the reference's ModelSystem pipeline, the CPU
crop/rotate glue and the Trainer are callers of this path, not part of it.
"""

import numpy as np

from .nn.help_func import make_list_if_not
from .nn.layers import (Concat, Conv2DToBatchedFixedWidthed, Convolutional2D, Flatten,
                        FullyConnected, LeakyRelu, Sigmoid, Upsample2D)
from .nn.losses import SegmentationDice2D, SoftmaxCrossEntropy
from .nn.models import Model
from .nn.optimizers import Adam
from .nn.regularizations import L2

CHAR_INPUT_HEIGHT = 32       # my_model/model.py:22
CHAR_FIXED_WIDTH = 8         # my_model/model.py:23
N_CHARS = 162                # len(primitives.CHARS), primitives/__init__.py:13-50
# output channels per sub-model = len(LAYER_NAMES[...]), my_model/constants.py
OUT_CHANNELS = {'monochrome': 1, 'paragraph': 1, 'line': 2}


def make_divisible_by(arr, y, x):
    """Zero-pads H and W up to the next multiple of (y, x) -- adding a FULL y / x when already
    aligned, like the reference (my_model/model.py:26-34).  Host-side (NumPy in, NumPy out)."""
    b, h, w, c = arr.shape
    add_y, add_x = y - h % y, x - w % x
    out = np.zeros((b, h + add_y, w + add_x, c), dtype=arr.dtype)
    out[:, add_y // 2:add_y // 2 + h, add_x // 2:add_x // 2 + w, :] = arr
    return out


def make_conv(out_ch, kernel_size=(5, 5), padding=2, **kwargs):
    return Convolutional2D(kernel_size, out_channels=out_ch, padding=padding,
                           regularizer=L2(0.01), **kwargs)


def make_conv_block(out_chs, last_sigmoid=False, **kwargs):
    out_chs = make_list_if_not(out_chs)
    layers, relations, prev = {}, {}, 0
    for i, out_ch in enumerate(out_chs, start=1):
        conv_name = f'conv_{i}'
        layers[conv_name] = make_conv(out_ch, **kwargs)
        if i == len(out_chs) and last_sigmoid is True:
            act_name, act = 'sigmoid', Sigmoid()
        else:
            act_name, act = f'leaky_relu_{i}', LeakyRelu(0.01)
        layers[act_name] = act
        relations[conv_name] = prev
        relations[act_name] = conv_name
        prev = act_name
    relations[0] = prev
    return Model(layers, relations)


def make_up(out_chs, **kwargs):
    return Model(layers={'upsample': Upsample2D(2), 'concat': Concat(),
                         'conv_block': make_conv_block(out_chs, **kwargs)},
                 relations={'upsample': 1, 'concat': ['upsample', 0], 'conv_block': 'concat',
                            0: 'conv_block'})


def make_single_up(out_chs, **kwargs):
    return Model(layers={'upsample': Upsample2D(2), 'conv_block': make_conv_block(out_chs, **kwargs)},
                 relations={'upsample': 0, 'conv_block': 'upsample', 0: 'conv_block'})


def wrap(name, model, **kwargs):
    return Model(layers={name: model}, relations={name: 0, 0: name}, **kwargs)


def make_monochrome(input_shape, optimizer=None):
    optimizer = Adam(lr=1e-2) if optimizer is None else optimizer
    kwargs = {'optimizer': optimizer, 'trainable': True}
    model = Model(layers={'Monochrome': make_conv_block([16, OUT_CHANNELS['monochrome']],
                                                        last_sigmoid=True, kernel_size=(3, 3),
                                                        padding=1, **kwargs)},
                  relations={'Monochrome': 0, 0: 'Monochrome'}, loss=SegmentationDice2D())
    model.initialize(input_shape)
    return model


def _make_hourglass(name, input_shape, channels, out_channels, optimizer):
    """make_paragraph / make_line share one topology (my_model/model.py:137-248): two stride-2
    conv blocks down, two (upsample x2 + conv block) up, a sigmoid conv block at the end."""
    optimizer = Adam(lr=1e-2) if optimizer is None else optimizer
    kwargs = {'optimizer': optimizer, 'trainable': True}
    downs, ups = [None, [channels], [channels]], [None, [channels], [channels]]
    layers = {
        **{f'down_{i}': make_conv_block(downs[i], kernel_size=(5, 5), padding=2, stride=2, **kwargs)
           for i in range(1, len(downs))},
        **{f'up_{i}': make_single_up(ups[i], kernel_size=(5, 5), padding=2, **kwargs)
           for i in range(1, len(ups))},
        'end': make_conv_block([out_channels], last_sigmoid=True, kernel_size=(5, 5), padding=2,
                               **kwargs),
    }
    relations = {
        'down_1': 0,
        **{f'down_{i + 1}': f'down_{i}' for i in range(1, len(downs) - 1)},
        f'up_{len(ups) - 1}': f'down_{len(downs) - 1}',
        **{f'up_{i}': f'up_{i + 1}' for i in range(1, len(ups) - 1)},
        'end': 'up_1',
        0: 'end',
    }
    model = wrap(name, Model(layers=layers, relations=relations), loss=SegmentationDice2D())
    model.initialize(input_shape)       # Paragraph (1 -> 1 channels): Model recognises the hourglass and attaches
    return model                        # its whole-network inference kernel (nn.models.HourglassFusion.detect)


def make_paragraph(input_shape, optimizer=None):
    return _make_hourglass('Paragraph', input_shape, 1, OUT_CHANNELS['paragraph'], optimizer)


def make_line(input_shape, optimizer=None):
    return _make_hourglass('Line', input_shape, 4, OUT_CHANNELS['line'], optimizer)


def make_dense_block(out_counts, **kwargs):
    out_counts = make_list_if_not(out_counts)
    layers, relations, prev = {}, {}, 0
    for i, n_out in enumerate(out_counts, start=1):
        dense_name = f'dense_{i}'
        layers[dense_name] = FullyConnected(n_output=n_out, **kwargs)
        relations[dense_name] = prev
        prev = dense_name
        if i < len(out_counts):
            act_name = f'leaky_relu_{i}'
            layers[act_name] = LeakyRelu(0.01)
            relations[act_name] = dense_name
            prev = act_name
    relations[0] = prev
    return Model(layers, relations)


def make_char(input_shape, optimizer=None):
    batch_size, _, width, in_channels = input_shape
    optimizer = Adam(lr=1e-2) if optimizer is None else optimizer
    kwargs = {'optimizer': optimizer, 'trainable': True}
    layers = {
        'conv_block': make_conv_block([64, 64, 64], kernel_size=(5, 3), padding=(0, 1),
                                      stride=(2, 1), **kwargs),
        'fixed_width': Conv2DToBatchedFixedWidthed(CHAR_FIXED_WIDTH),
        'flatten': Flatten(),
        'dense_block': make_dense_block([1024, 128, N_CHARS], **kwargs),
    }
    relations = {'conv_block': 0, 'fixed_width': 'conv_block', 'flatten': 'fixed_width',
                 'dense_block': 'flatten', 0: 'dense_block'}
    model = wrap('Char', Model(layers=layers, relations=relations), loss=SoftmaxCrossEntropy())
    model.initialize((batch_size, CHAR_INPUT_HEIGHT, width, in_channels))
    return model


MAKERS = {'monochrome': make_monochrome, 'paragraph': make_paragraph, 'line': make_line,
          'char': make_char}


# ---- model_weights.json (my_model/train.py:132-141, my_model/predict.py:12-23) -----------------
# JSON stays the source of truth; `weights_io` keeps a hash-checked binary sidecar next to it.
from .weights_io import load_weights, save_weights  # noqa: E402,F401
