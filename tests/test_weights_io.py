"""model_weights.json sidecar cache: the JSON written is byte-identical to
what the reference's `json.dump(weights, fp, separators=(',', ':'))` (my_model/train.py:141)
produces, the sidecar is used only while it matches the JSON's hash, and merge semantics follow
`update_weights_func` (train.py:132-141)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from univer_ocr_b200 import weights_io  # noqa: E402


class Net:
    def __init__(self, prefix, seed):
        rng = np.random.default_rng(seed)
        self.w = {f'{prefix}/conv_1': {'w': rng.normal(size=(3, 3, 1, 4)).astype(np.float32).astype(np.float64),
                                       'b': rng.normal(size=(4,)).astype(np.float32).astype(np.float64)},
                  f'{prefix}/block/dense_1': {'w': rng.normal(size=(5, 2)).astype(np.float32).astype(np.float64)}}

    def get_weights(self):
        return {layer: {n: v.tolist() for n, v in params.items()} for layer, params in self.w.items()}

    def set_weights(self, weights):
        for layer, params in self.w.items():
            for n in params:
                new = weights.get(layer, {}).get(n)
                if new is not None:
                    params[n] = np.array(new)


def test_json_bytes_are_the_reference_format(tmp_path):
    path = tmp_path / 'model_weights.json'
    a = Net('Monochrome', 1)
    weights_io.save_weights(a, path)
    want = json.dumps(a.get_weights(), separators=(',', ':')).encode()
    assert path.read_bytes() == want
    assert os.path.exists(weights_io.sidecar_path(path))


def test_sidecar_hit_equals_json_parse_and_detects_stale(tmp_path):
    path = tmp_path / 'model_weights.json'
    a, b = Net('Monochrome', 1), Net('Char', 2)
    weights_io.save_weights([a, b], path)
    stats = {}
    got = weights_io.read(path, stats=stats)
    assert stats['source'] == 'sidecar'
    plain = json.loads(path.read_bytes())
    assert set(got) == set(plain)
    for layer in plain:
        for n in plain[layer]:
            assert got[layer][n].dtype == np.float64
            assert np.array_equal(got[layer][n], np.array(plain[layer][n]))
    # someone (the reference) rewrites the JSON: the sidecar no longer matches and is rebuilt
    plain['Monochrome/conv_1']['b'] = [9.0, 8.0, 7.0, 6.0]
    path.write_text(json.dumps(plain, separators=(',', ':')))
    stats = {}
    got = weights_io.read(path, stats=stats)
    assert stats['source'] == 'json' and got['Monochrome/conv_1']['b'].tolist() == [9.0, 8.0, 7.0, 6.0]
    stats = {}
    weights_io.read(path, stats=stats)
    assert stats['source'] == 'sidecar'
    # a truncated sidecar is ignored, not trusted
    with open(weights_io.sidecar_path(path), 'r+b') as fp:
        fp.truncate(40)
    stats = {}
    assert weights_io.read(path, stats=stats)['Monochrome/conv_1']['b'].tolist() == [9.0, 8.0, 7.0, 6.0]
    assert stats['source'] == 'json'


def test_merge_keeps_other_models_and_load_restores(tmp_path, capsys):
    path = tmp_path / 'model_weights.json'
    a, b = Net('Monochrome', 1), Net('Char', 2)
    weights_io.save_weights(a, path)
    weights_io.save_weights(b, path)                           # must not drop Monochrome
    a2, b2 = Net('Monochrome', 3), Net('Char', 4)
    weights_io.load_weights([a2, b2], path)
    for fresh, old in ((a2, a), (b2, b)):
        for layer in old.w:
            for n in old.w[layer]:
                assert np.array_equal(fresh.w[layer][n], old.w[layer][n])
    assert weights_io.load_weights(Net('Line', 5), tmp_path / 'absent.json') == {}
    assert 'No model_weights.json file found' in capsys.readouterr().out


def test_cache_off_never_touches_a_sidecar(tmp_path):
    path = tmp_path / 'w.json'
    weights_io.save_weights(Net('Line', 7), path, cache=False)
    assert not os.path.exists(weights_io.sidecar_path(path))
    stats = {}
    weights_io.read(path, cache=False, stats=stats)
    assert stats['source'] == 'json' and not os.path.exists(weights_io.sidecar_path(path))
