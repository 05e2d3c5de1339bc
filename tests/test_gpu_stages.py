"""GPU parity tests of the crop stages (univer_ocr_b200/stages.py, csrc/stages.cu) against
scipy.ndimage itself and against the oracle restatement of the reference's stage logic (oracle/np_stages.py, pinned
against the reference's functions in tests/test_oracle_pin.py).  Everything here is selection / resampling with SciPy's
own arithmetic: bit-exact."""
import ctypes

import numpy as np
import pytest

from oracle import np_stages as S
from tests import stage_cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nn():
    import univer_ocr_b200.nn as nn_
    nn_.CP.use_gpu()
    return nn_


def test_zoom_nearest_matches_scipy(nn):
    """uocr_zoom_nearest_f32 == ndimage.zoom(a, (1, zf, zf, 1), order=0) (+ the zero padding to minimal_width) through
    the C ABI: random shapes and factors, incl. the rounding case where the last coordinate lands above in - 1."""
    from scipy import ndimage
    from univer_ocr_b200 import stages
    rng = np.random.default_rng(3)
    cases = [((1, 21, 197, 2), 16 / 21), ((1, 42, 256, 1), 64 / 42), ((2, 15, 184, 3), 32 / 15), ((1, 1, 1, 1), 3.0),
             ((1, 7, 5, 1), 1.0)]
    for t in range(40):
        h, w, c = int(rng.integers(1, 70)), int(rng.integers(1, 300)), int(rng.integers(1, 4))
        cases.append(((int(rng.integers(1, 3)), h, w, c), int(rng.integers(1, 64)) / h))
    for shape, zf in cases:
        a = rng.uniform(-1, 1, size=shape).astype(np.float32)
        want = ndimage.zoom(a, (1, zf, zf, 1), order=0)
        got = stages.zoom_nearest(nn.CP.copy(a), zf).get()
        assert got.shape == want.shape and np.array_equal(got, want), (shape, zf)
        padded = stages.zoom_nearest(nn.CP.copy(a), zf, minimal_width=want.shape[2] + 9).get()
        assert np.array_equal(padded[:, :, :want.shape[2]], want) and not padded[:, :, want.shape[2]:].any()


def test_rotate_matches_scipy(nn):
    """uocr_rotate_f32 (order 0 and 1) and uocr_rotate_nearest_u8 == ndimage.rotate(a, angle, axes=(2, 1), order,
    reshape=True): random angles, the quarter turns (pure permutations) and 45 degrees (coordinates exactly on .5)."""
    from scipy import ndimage
    from univer_ocr_b200 import stages
    rng = np.random.default_rng(4)
    for t in range(60):
        n, h, w, c = int(rng.integers(1, 3)), int(rng.integers(1, 60)), int(rng.integers(1, 90)), int(rng.integers(1, 3))
        a = rng.uniform(-1, 1, size=(n, h, w, c)).astype(np.float32)
        angle = float(rng.uniform(0, 180)) if t % 3 else float(rng.choice([0, 90, 180, 270, 45, 135, 30]))
        for order in (0, 1):
            want = ndimage.rotate(a, angle, axes=(2, 1), order=order, reshape=True)
            got = stages.rotate_array(nn.CP.copy(a), angle, good_rotation=bool(order)).get()
            assert got.shape == want.shape and np.array_equal(got, want), (a.shape, angle, order)
        m = a[..., :1] > 0
        want = ndimage.rotate(m, angle, axes=(2, 1), order=0, reshape=True)
        got = stages.rotate_array(stages._device(m), angle, good_rotation=False).get()
        assert np.array_equal(got.astype(bool), want)
        # the rows the rotated mask spans, read off without materialising it (the angle search's probe)
        if want.any():
            rows = np.flatnonzero(want.any(axis=(0, 2, 3)))
            other = float(rng.uniform(0, 180))
            want2 = ndimage.rotate(m, other, axes=(2, 1), order=0, reshape=True)
            if want2.any():
                rows2 = np.flatnonzero(want2.any(axis=(0, 2, 3)))
                assert stages.rotated_heights(stages._device(m), (angle, other)) == [rows[-1] - rows[0] + 1,
                                                                                     rows2[-1] - rows2[0] + 1]
            assert stages.rotated_heights(stages._device(m), (angle,)) == [rows[-1] - rows[0] + 1]
            assert stages.rotated_height(stages._device(m), angle) == rows[-1] - rows[0] + 1
    assert stages.rotate_array('untouched', None) == 'untouched'                  # angle None: the array itself


def test_crop_bbox_and_above_mean(nn):
    """uocr_crop_masked_f32 == (image * mask)[:, ry, rx, :] incl. the -0.0 a product leaves under a False mask,
    uocr_crop_label_mask, uocr_mask_bbox == ndimage.find_objects of the mask (IndexError when empty),
    uocr_above_mean_mask == arr > mean(arr), uocr_channel_slice_u8."""
    from scipy import ndimage
    from univer_ocr_b200 import glue, stages
    rng = np.random.default_rng(6)
    pred = (rng.uniform(size=(1, 60, 90, 1)) ** 3).astype(np.float32)
    fg = pred.astype(np.float64) > pred.astype(np.float64).mean()
    assert np.array_equal(glue.above_mean(pred).get().astype(bool), fg)
    labels, objects = stages.label_objects(pred)
    want_labels, count = ndimage.label(fg)
    assert len(objects) == count and np.array_equal(labels.get(), want_labels)
    image = rng.uniform(-1, 1, size=(1, 60, 90, 3)).astype(np.float32)
    dimage = nn.CP.copy(image)
    for l in (1, count // 2 + 1, count):
        mask = want_labels == l
        ry, rx = objects[l - 1]['slices']
        assert (ry, rx) == ndimage.find_objects(mask.astype(np.uint8))[0][1:3]
        want = (image * mask)[:, ry, rx, :]
        got = stages.crop(dimage, ry, rx, labels, l).get()
        assert np.array_equal(got, want) and np.array_equal(np.signbit(got), np.signbit(want))
        assert np.array_equal(stages.crop(dimage, ry, rx).get(), image[:, ry, rx, :])
        cm = stages.crop_label_mask(labels, l, ry, rx)
        assert np.array_equal(cm.get().astype(bool), mask[:, ry, rx, :])
        assert stages.mask_bbox(cm) == (slice(0, ry.stop - ry.start), slice(0, rx.stop - rx.start))
    whole = stages._device(want_labels == 1)
    assert stages.mask_bbox(whole) == ndimage.find_objects((want_labels == 1).astype(np.uint8))[0][1:3]
    with pytest.raises(IndexError):
        stages.mask_bbox(stages._device(np.zeros((1, 5, 7, 1), bool)))
    two = (rng.uniform(size=(2, 9, 11, 2)) < 0.5).astype(np.uint8)
    for k in (0, 1):
        assert np.array_equal(glue.channel(stages._device(two), k).get(), two[..., k:k + 1])
    planes = glue.channel_planes(stages._device(two[:1])).get()
    assert planes.shape == (2, 9, 11, 1) and np.array_equal(planes[:, :, :, 0], np.moveaxis(two[0], -1, 0))
    # error convention of the C ABI: a region outside the image
    from univer_ocr_b200._lib import UocrError, lib
    out = nn.DeviceArray.empty((1, 4, 4, 3))
    with pytest.raises(UocrError) as err:
        lib.uocr_crop_masked_f32(dimage.ptr, None, 0, out.ptr, 1, 60, 90, 3, 58, 0, 4, 4, nn.CP.stream())
    assert err.value.code == -1 and 'region outside' in str(err.value)


def test_crop_rotate_and_zoom_lines_matches_oracle(nn):
    """stages.CropRotateAndZoomLines (interpreter.py:423-523) == the oracle's serial restatement, for the four reading
    directions rearrange_lines distinguishes, with and without zoom / padding: same nesting, same shapes, same bits."""
    from univer_ocr_b200 import stages
    for direction in (None, 90, 180, 270):
        m1, a1 = C.line_paragraph(1, direction)
        m2, a2 = C.line_paragraph(2, direction, lines=2)
        masks, arrays = [m1, m2], [[a1[0], a2[0]], [a1[1], a2[1]]]
        for zoomed, minimal in ((32, 200), (32, 1000), (None, 300), (None, None)):
            want = S.crop_rotate_and_zoom_lines(masks, arrays, zoomed, minimal)
            got = stages.CropRotateAndZoomLines(8, zoomed, minimal, to_host=True)(masks, arrays)
            assert len(got) == len(want) == 2
            for aid in range(2):
                assert [len(p) for p in got[aid]] == [len(p) for p in want[aid]] == [3, 2]
                for pid in range(2):
                    for g, w in zip(got[aid][pid], want[aid][pid]):
                        assert g.shape == w.shape and np.array_equal(g, w), (direction, zoomed, minimal, aid, pid)
    # device results by default
    got = stages.CropRotateAndZoomLines(None, 32, 200)(masks, arrays)
    assert isinstance(got[0][0][0], nn.DeviceArray) and got[0][0][0].shape[1] == 32
    # the reference's failure modes: no marks -> IndexError; no reading direction -> UnboundLocalError
    with pytest.raises(IndexError):
        stages.CropRotateAndZoomLines(None, 32, 200)([np.zeros((1, 16, 16, 2), np.float32)], [[a1[0][:, :16, :16]]])
    same = np.zeros((1, 32, 32, 2), np.float32)
    same[0, 10:12, 4:28, :] = 1.0                             # top and bottom marks coincide: offset 0 in both axes
    with pytest.raises(UnboundLocalError):
        stages.CropRotateAndZoomLines(None, 32, 200)([same], [[a1[0][:, :32, :32]]])
    with pytest.raises(UnboundLocalError):
        S.crop_rotate_and_zoom_lines([same], [[a1[0][:, :32, :32]]], 32, 200)


def test_crop_and_rotate_paragraphs_matches_oracle(nn):
    """stages.CropAndRotateParagraphs (interpreter.py:234-374) == the oracle's serial restatement: the same paragraphs,
    the same angles out of the ternary search over nearest-rotated masks, the same straightened crops, bit for bit --
    with the rotation search and without it."""
    from univer_ocr_b200 import stages
    for seed, tilt in ((0, (12.0, -25.0)), (1, (80.0, 3.0))):
        pred, images = C.paragraph_page(seed, tilt=tilt)
        for find_rotation in (True, False):
            want, angles = S.crop_and_rotate_paragraphs(pred, images, find_rotation)
            stage = stages.CropAndRotateParagraphs(4, find_rotation, to_host=True)
            got = stage(pred, images)
            assert stage.angles == angles and len(angles) == 2
            assert len(got) == len(want) == 2
            for iid in range(2):
                for g, w in zip(got[iid], want[iid]):
                    assert g.shape == w.shape and np.array_equal(g, w), (seed, find_rotation, iid)


def test_text_pipeline_wiring_matches_oracle_composition(nn):
    """predict.TextPipeline = the reference's PREDICT model system (my_model/model.py:688-717) on the device.  The four
    networks are replaced by deterministic stand-ins (so that the comparison is about the wiring, the padding between
    the stages and the stages themselves, not about TF32 arithmetic at decision thresholds): the same stand-ins composed
    on the host with the oracle's stages must give the same paragraphs, angles, line crops (bit for bit), class ids and
    text."""
    from oracle import np_oracle as O
    from univer_ocr_b200 import my_model, predict
    pred, images = C.paragraph_page(3, h=150, w=230)
    page = images[0]

    def lines_for(shape):                                      # Line stand-in: marks at fixed relative rows
        _, h, w, _ = shape
        out = np.zeros((1, h, w, 2), np.float32)
        for centre in (h // 3, (2 * h) // 3):
            out[0, centre - 6:centre - 4, 6:w - 6, 0] = 1.0
            out[0, centre + 4:centre + 6, 6:w - 6, 1] = 1.0
        return out

    def chars_for(line):                                       # Char stand-in: class from the column's brightness
        cols = np.asarray(line)[0, :, :, 0].mean(axis=0)
        out = np.zeros((cols.size, my_model.N_CHARS), np.float32)
        out[np.arange(cols.size), (cols * 40).astype(int) % my_model.N_CHARS] = 1.0
        return out

    padded_pred = O.make_divisible_by(pred, 16, 16).astype(np.float32)
    predictors = {'monochrome': lambda x: x,
                  'paragraph': lambda x: nn.CP.copy(padded_pred),
                  'line': lambda x: nn.CP.copy(lines_for(x.shape)),
                  'char': lambda x: nn.CP.copy(chars_for(x.get()))}
    chars = [chr(33 + i) for i in range(my_model.N_CHARS)]
    similar = lambda a, b: a == b                              # noqa: E731
    got = predict.TextPipeline(predictors=predictors, chars=chars, are_similar=similar)(page)
    # the same composition on the host
    x = O.make_divisible_by(page, 16, 16).astype(np.float32)
    cropped, angles = S.crop_and_rotate_paragraphs(padded_pred, [x], True)
    cropped = [O.make_divisible_by(t, 16, 16).astype(np.float32) for t in cropped[0]]
    line_pred = [lines_for(t.shape) for t in cropped]
    lines = S.crop_rotate_and_zoom_lines(line_pred, [cropped], my_model.CHAR_INPUT_HEIGHT, my_model.CHAR_FIXED_WIDTH)[0]
    assert got['angles'] == angles and len(angles) == 2
    assert len(got['cropped_monochrome']) == len(cropped)
    for g, w in zip(got['cropped_monochrome'], cropped):
        assert np.array_equal(g.get(), w)
    assert [len(p) for p in got['cropped_2_monochrome']] == [len(p) for p in lines] == [2, 2]
    for gp, wp, tp in zip(got['cropped_2_monochrome'], lines, got['text']):
        for g, w, text in zip(gp, wp, tp):
            assert g.shape == w.shape and w.shape[1] == my_model.CHAR_INPUT_HEIGHT and np.array_equal(g.get(), w)
            assert text == O.pred_to_text(chars_for(w), chars, similar) and len(text) > 0
    ids = predict.TextPipeline(predictors=predictors)(page)['text']
    assert all(isinstance(v, int) for p in ids for line in p for v in line)


def test_text_pipeline_runs_the_networks(nn):
    """The same pipeline with the real networks (default initialisation, TF32 mode) on a small page: every stage
    runs, shapes chain (crops padded to multiples of 16, lines zoomed to CHAR_INPUT_HEIGHT), the text nesting follows
    the paragraphs and lines found.  find_rotation off: with untrained weights the paragraph map is noise and the
    search would only slow the test."""
    from univer_ocr_b200 import my_model, predict
    rng = np.random.default_rng(8)
    page = rng.uniform(0, 1, size=(1, 90, 120, 1)).astype(np.float32)
    pipe = predict.TextPipeline(find_rotation=False)
    mono = pipe.predict('monochrome', nn.CP.copy(np.zeros((1, 96, 128, 1), np.float32)))
    assert mono.shape == (1, 96, 128, 1)
    # untrained networks label noise: cap the work by handing the Paragraph stage a clean synthetic map
    clean = np.zeros((1, 96, 128, 1), np.float32)
    clean[0, 20:70, 16:110, 0] = 1.0
    pipe.predictors['paragraph'] = lambda x: nn.CP.copy(clean)
    marks = np.zeros((1, 64, 96, 2), np.float32)
    marks[0, 20:22, 8:88, 0] = 1.0
    marks[0, 40:42, 8:88, 1] = 1.0
    pipe.predictors['line'] = lambda x: nn.CP.copy(marks[:, :x.shape[1], :x.shape[2]]) if x.shape[1:3] == (64, 96) else None
    out = pipe(page)
    assert len(out['cropped_monochrome']) == 1 and out['cropped_monochrome'][0].shape == (1, 64, 96, 1)
    assert len(out['cropped_2_monochrome'][0]) == 1
    line = out['cropped_2_monochrome'][0][0]
    assert line.shape[1] == my_model.CHAR_INPUT_HEIGHT and line.shape[2] >= my_model.CHAR_FIXED_WIDTH
    pred = out['char_pred'][0][0]
    assert pred.shape == (line.shape[2], my_model.N_CHARS) and np.isfinite(pred.get()).all()
    assert isinstance(out['text'][0][0], list)


def test_stages_with_nothing_to_cut(nn):
    """A constant Paragraph map has no pixel above its mean: no objects, empty result lists (one per image), like the
    reference's label_layer -> {} -> [[] for image in images]; the pipeline then returns no text."""
    from univer_ocr_b200 import predict, stages
    flat = np.full((1, 48, 64, 1), 0.25, np.float32)
    images = [np.ones((1, 48, 64, 1), np.float32), np.ones((1, 48, 64, 2), np.float32)]
    assert stages.CropAndRotateParagraphs()(flat, images) == [[], []]
    want, angles = S.crop_and_rotate_paragraphs(flat, images, True)
    assert want == [[], []] and angles == []
    assert stages.CropRotateAndZoomLines(None, 32, 8)([], [[], []]) == [[], []]
    pipe = predict.TextPipeline(predictors={'monochrome': lambda x: x, 'paragraph': lambda x: nn.CP.copy(flat)})
    out = pipe(np.zeros((1, 40, 60, 1), np.float32))
    assert out['text'] == [] and out['angles'] == []


def test_stages_match_reference_golden(nn):
    """The device stage classes against tests/golden/stages.npz, i.e. against what the UNMODIFIED reference's own
    functions returned for the same seeded inputs (tests/golden/make_stage_golden.py): zoomed line crops in all four
    reading directions, paragraph angles, straightened paragraph crops -- bit for bit."""
    import os
    from univer_ocr_b200 import stages
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'stages.npz'))
    for direction in (None, 90, 180, 270):
        mask, arrays = C.line_paragraph(1, direction)
        tag = f'lines_{direction}'
        got = stages.CropRotateAndZoomLines(8, 32, 200, to_host=True)([mask], [[a] for a in arrays])
        assert len(got[0][0]) == int(gold[f'{tag}__count'])
        for lid in range(len(got[0][0])):
            for aid in range(2):
                want = gold[f'{tag}__line{lid}_array{aid}']
                assert got[aid][0][lid].shape == want.shape and np.array_equal(got[aid][0][lid], want)
    for seed, tilt in ((0, (12.0, -25.0)), (1, (80.0, 3.0))):
        pred, images = C.paragraph_page(seed, tilt=tilt)
        tag = f'page_{seed}'
        stage = stages.CropAndRotateParagraphs(4, True, to_host=True)
        got = stage(pred, images)
        assert len(stage.angles) == int(gold[f'{tag}__count'])
        for pid, angle in enumerate(stage.angles):
            assert angle == float(gold[f'{tag}__angle{pid}'])
            for iid in range(2):
                want = gold[f'{tag}__par{pid}_image{iid}']
                assert got[iid][pid].shape == want.shape and np.array_equal(got[iid][pid], want)
