"""Route A of INTEGRATION.md executed literally (missing #5): the reference's UNMODIFIED
`my_model/model.py:108-304` builders (and its `nn/model_system.py`) construct their networks out of
`univer_ocr_b200.nn` -- `..nn` is aliased to the drop-in package by `oracle/ref_loader.load_my_model_on` -- and the
result is compared with what the repo's own mirror `univer_ocr_b200.my_model` builds: flattened layer names and
order, relations, layer hyper-parameters, every intermediate output shape, parameter counts, fusion plans, the
whole-network inference kernel and the fused-update eligibility.  Construction and shape analysis need no device
(parameters are uploaded on first use), so this runs in the authoring container; it needs /root/reference and is
skipped on the GPU box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='needs /root/reference')

SHAPES = {'monochrome': (2, 496, 736, 1), 'paragraph': (2, 496, 736, 1), 'line': (2, 128, 256, 1),
          'char': (2, 32, 256, 1)}


@pytest.fixture(scope='module')
def ref_on_b200():
    import univer_ocr_b200.nn as nn
    import univer_ocr_b200.nn.progress_tracker  # noqa: F401  (my_model/model.py imports track_function from it)
    return ref_loader.load_my_model_on(nn)


def _layer_signature(layer):
    keys = ('kernel_size', 'in_channels', 'out_channels', 'padding', 'padding_value', 'stride', 'bias', 'alpha',
            'scale_factor', 'width', 'n_input', 'n_output', 'axis', 'trainable')
    sig = {k: getattr(layer, k) for k in keys if hasattr(layer, k)}
    sig['class'] = type(layer).__name__
    sig['regularizer'] = repr(getattr(layer, 'regularizer', None))
    sig['params'] = {k: tuple(p.shape) for k, p in layer.params().items()}
    return sig


@pytest.mark.parametrize('name', list(SHAPES))
def test_reference_builders_construct_the_same_networks_on_the_dropin_package(ref_on_b200, name):
    import univer_ocr_b200.nn as nn
    from univer_ocr_b200 import my_model
    shape = SHAPES[name]
    theirs = getattr(ref_on_b200, f'make_{name}')(shape, nn.optimizers.Adam(lr=0.0015))
    ours = my_model.MAKERS[name](shape, nn.optimizers.Adam(lr=0.0015))
    # the reference's code built drop-in objects, not its own
    assert type(theirs).__module__ == 'univer_ocr_b200.nn.models'
    assert all(type(l).__module__.startswith('univer_ocr_b200.nn') for l in theirs.layers.values())
    assert list(theirs.layers) == list(ours.layers)                       # flattened names = model_weights.json keys
    assert theirs.relations == ours.relations
    for lname in ours.layers:
        assert _layer_signature(theirs.layers[lname]) == _layer_signature(ours.layers[lname]), lname
    assert theirs.get_all_output_shapes([shape]) == ours.get_all_output_shapes([shape])
    assert theirs.count_parameters() == ours.count_parameters()
    assert type(theirs.loss) is type(ours.loss)
    # the fast paths are found from the topology, not from builder hints
    assert theirs._plan_train == ours._plan_train and theirs._plan_infer == ours._plan_infer
    assert any(step[0] != 'layer' for step in theirs._plan_infer)
    assert type(theirs.infer_fusion) is type(ours.infer_fusion)
    assert (theirs.infer_fusion is not None) == (name in ('paragraph', 'line'))    # the two one-kernel hourglass paths
    assert theirs.fused_optimizer() is not None                           # Model.train takes the flat fused update
    if name in ('paragraph', 'line', 'monochrome'):
        assert theirs.get_receptive_fields() == ours.get_receptive_fields()


def test_reference_model_system_assembles_on_the_dropin_package(ref_on_b200):
    """`make_model_system` (my_model/model.py:486-717) for the single-network training modes: the reference's
    ModelSystem / ModelComponent (its own nn/model_system.py, unmodified) wrap drop-in Models."""
    import univer_ocr_b200.nn as nn
    mm = ref_on_b200
    assert mm.ModelSystem.__module__.endswith('nn.model_system') and 'univer_ocr_b200' not in mm.ModelSystem.__module__
    weights = {'Monochrome/conv_1': {'b': np.linspace(-1, 1, 16).tolist()}}
    for mode, model_name, shape in ((mm.Modes.TRAIN_MONOCHROME, 'Monochrome', (1, 32, 48, 1)),
                                    (mm.Modes.TRAIN_PARAGRAPH, 'Paragraph', (1, 32, 48, 1))):
        system, models, names = mm.make_model_system(shape, optimizer=nn.optimizers.Adam(lr=0.001), weights=weights,
                                                     mode=mode)
        assert names == [model_name] and list(models) == [model_name]
        assert type(models[model_name]).__module__ == 'univer_ocr_b200.nn.models'
        assert isinstance(system, mm.ModelSystem)
    # set_weights went through the drop-in's soft-fail loader without a device: pending host values
    mono = mm.make_model_system((1, 32, 48, 1), weights=weights, mode=mm.Modes.TRAIN_MONOCHROME)[1]['Monochrome']
    b = mono.layers['Monochrome/conv_1'].b
    assert b._pending is not None and np.allclose(b._pending, np.linspace(-1, 1, 16))
