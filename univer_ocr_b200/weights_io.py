"""`model_weights.json` with a binary sidecar cache.

The reference keeps all weights of the four networks in one text file -- 803 395 numbers as
float64 repr text, 16.5 MB -- that `my_model/train.py:132-141` re-reads, merges and re-writes
whenever a model improves and that `my_model/train.py:119-123` / `my_model/predict.py:12-23` parse
on every start.  The JSON file stays the source of truth here, byte-compatible with the
reference (compact separators, nested lists, `{layer_name: {param_name: [...]}}`).  Next to it
lives `<file>.npz`: the same tensors as float64 arrays plus the SHA-256 of the JSON bytes they
were parsed from.  A reader hashes the JSON (milliseconds), and if the sidecar carries that hash
it loads the arrays instead of parsing the text; a stale or missing sidecar is rebuilt from the
JSON, never trusted.  Writers replace both files atomically (temp file + rename), JSON first.

Models only need the reference's `get_weights() / set_weights(dict)` protocol
(`nn/layers/layers.py:120-137`), so this module has no device code of its own.
"""
import hashlib
import json
import os
import tempfile
import zipfile

import numpy as np

_HASH_KEY = '__json_sha256__'
_SEP = '//'                                   # layer names contain '/', param names do not


def sidecar_path(path):
    return str(path) + '.npz'


def _sha256(data):
    return hashlib.sha256(data).hexdigest()


def _atomic_write(path, write):
    directory = os.path.dirname(os.path.abspath(path))
    fd, tmp = tempfile.mkstemp(dir=directory, prefix='.' + os.path.basename(path) + '.')
    try:
        with os.fdopen(fd, 'wb') as fp:
            write(fp)
        os.replace(tmp, path)
    except BaseException:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise


def _to_arrays(weights):
    """{layer: {param: nested list}} → same with float64 arrays; anything that is not a
    rectangular numeric tensor (hand-edited files) stays as it is."""
    out = {}
    for layer, params in weights.items():
        out[layer] = {}
        for name, value in params.items():
            try:
                arr = np.asarray(value, dtype=np.float64)
            except (ValueError, TypeError):
                arr = value
            out[layer][name] = arr
    return out


def _write_sidecar(path, weights, digest):
    flat = {_HASH_KEY: np.frombuffer(digest.encode(), dtype=np.uint8)}
    for layer, params in weights.items():
        for name, value in params.items():
            if not isinstance(value, np.ndarray):
                return False                                  # not cacheable; JSON path only
            flat[layer + _SEP + name] = value
    _atomic_write(sidecar_path(path), lambda fp: np.savez(fp, **flat))
    return True


def _read_sidecar(path, digest):
    try:
        with np.load(sidecar_path(path)) as data:
            if _HASH_KEY not in data.files or bytes(data[_HASH_KEY]).decode() != digest:
                return None
            weights = {}
            for key in data.files:
                if key == _HASH_KEY:
                    continue
                layer, name = key.rsplit(_SEP, 1)
                weights.setdefault(layer, {})[name] = data[key]
            return weights
    except (OSError, ValueError, KeyError, EOFError, zipfile.BadZipFile):
        return None


def read(path, cache=True, stats=None):
    """→ {layer_name: {param_name: float64 array}} of the JSON file at `path`; {} when the file
    does not exist (the reference prints 'No model_weights.json file found' and starts from the
    initialiser, `train.py:121-123`).  `stats['source']` reports 'sidecar' / 'json' / 'missing'."""
    stats = {} if stats is None else stats
    try:
        with open(path, 'rb') as fp:
            raw = fp.read()
    except OSError:
        stats['source'] = 'missing'
        return {}
    digest = _sha256(raw)
    if cache:
        weights = _read_sidecar(path, digest)
        if weights is not None:
            stats['source'] = 'sidecar'
            return weights
    weights = _to_arrays(json.loads(raw))
    stats['source'] = 'json'
    if cache:
        try:
            _write_sidecar(path, weights, digest)
        except OSError:
            pass                                              # read-only location: no cache, no error
    return weights


def write(path, weights, cache=True):
    """Writes `weights` ({layer: {param: array or nested list}}) as the reference's JSON and
    refreshes the sidecar."""
    plain = {layer: {name: (v.tolist() if isinstance(v, np.ndarray) else v) for name, v in params.items()}
             for layer, params in weights.items()}
    raw = json.dumps(plain, separators=(',', ':')).encode()   # train.py:141
    _atomic_write(path, lambda fp: fp.write(raw))
    if cache:
        try:
            _write_sidecar(path, _to_arrays(plain), _sha256(raw))
        except OSError:
            pass


def load_weights(models, path, cache=True, stats=None):
    """`json.load` + `model.set_weights(weights)` for every model (`train.py:119-126`)."""
    stats = {} if stats is None else stats
    weights = read(path, cache=cache, stats=stats)
    if stats.get('source') == 'missing':
        print('No model_weights.json file found')
    for model in (models if isinstance(models, (list, tuple)) else [models]):
        model.set_weights(weights)
    return weights


def save_weights(models, path, cache=True):
    """Read-modify-write merge of the models' weights into the file (`train.py:132-141`): layers
    of other models already in the file are kept."""
    weights = read(path, cache=cache)
    for model in (models if isinstance(models, (list, tuple)) else [models]):
        weights.update(model.get_weights())
    write(path, weights, cache=cache)
    return weights
