"""Case tables shared by tests/golden/make_golden.py (which runs the reference) and the
parity tests (which run the oracle and the CUDA path on the same cases)."""

CONV_CASES = [
    # name, (n,h,w), cin, cout, ks, padding, padding_value, stride
    ('plain',        (3, 5, 5), 6, 7, (4, 4), 0, 0.0, 1),        # test_gradients.py:130-165
    ('pad',          (3, 5, 5), 6, 7, (4, 4), 1, 0.0, 1),
    ('padval',       (3, 5, 5), 6, 7, (4, 4), 1, 0.5, 1),
    ('stride',       (3, 5, 5), 6, 7, (4, 4), 0, 0.0, 2),
    ('pad_stride',   (3, 5, 5), 6, 7, (4, 4), 1, 0.0, 2),
    ('ident_padval', (2, 12, 16), 6, 7, (3, 3), 1, 0.5, 1),       # test_identity.py:9-26 shapes
    ('ident_ps',     (2, 12, 16), 6, 7, (3, 3), 1, 0.0, 2),
    ('mono1',        (2, 16, 24), 1, 16, (3, 3), 1, 0.0, 1),      # my_model/model.py:119-122
    ('mono2',        (2, 16, 24), 16, 1, (3, 3), 1, 0.0, 1),
    ('para_down',    (2, 16, 24), 1, 1, (5, 5), 2, 0.0, 2),       # :155-160
    ('para_odd',     (1, 13, 11), 1, 1, (5, 5), 2, 0.0, 2),       # ragged: (H+2p-k) % s != 0
    ('line_mid',     (2, 8, 16), 4, 4, (5, 5), 2, 0.0, 1),        # :212-223
    ('line_down',    (2, 16, 16), 4, 4, (5, 5), 2, 0.0, 2),
    ('line_end',     (2, 8, 16), 4, 2, (5, 5), 2, 0.0, 1),
    ('char1',        (2, 32, 10), 1, 64, (5, 3), (0, 1), 0.0, (2, 1)),   # :283-285
    ('char2',        (2, 14, 10), 64, 64, (5, 3), (0, 1), 0.0, (2, 1)),
    ('char3',        (2, 5, 10), 64, 64, (5, 3), (0, 1), 0.0, (2, 1)),
    ('rect_asym',    (1, 9, 7), 3, 5, (2, 3), (1, 0), -0.25, (1, 2)),
]

POOL_CASES = [
    # name, (n,h,w,c), k, padding, stride, ceil_mode
    ('k2',        (2, 6, 8, 3), 2, 0, None, False),               # test_identity.py:29-42
    ('k2_pad',    (2, 6, 8, 3), 2, 1, None, False),
    ('k2_s1',     (2, 6, 8, 3), 2, 0, 1, False),
    ('k2_pad_s1', (2, 6, 8, 3), 2, 1, 1, False),
    ('k3',        (2, 9, 7, 2), 3, 0, None, False),
    ('k3_ceil',   (2, 8, 7, 2), 3, 0, None, True),
    ('k2_ceil',   (1, 5, 7, 2), 2, 0, None, True),
    ('k32_s21',   (1, 7, 9, 2), (3, 2), (1, 0), (2, 1), False),
]

MODEL_SHAPES = {'monochrome': (2, 16, 32, 1), 'paragraph': (2, 16, 32, 1),
                'line': (2, 16, 32, 1), 'char': (2, 32, 12, 1)}
