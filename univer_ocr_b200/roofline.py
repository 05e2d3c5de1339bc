"""Algorithmic work per layer call -- the per-unit figures of SURVEY.md 8(d), stated in DESIGN.md:

  conv (each of fwd / dgrad / wgrad):  flops = 2 N Ho Wo kh kw Cin Cout
                                       bytes = 4 (N H W Cin + N Ho Wo Cout + kh kw Cin Cout)
  FC fwd:                              flops = 2 B (n_in + 1) n_out, bytes = 4 (B n_in + B n_out + W)
  elementwise fwd: 8 B/elem; bwd: 12 B/elem;  upsample fwd: 4 (1 + sy sx) B per input element
  window batch: 4 (1 + width) B per input element

Which roofline binds: tensor peak for the Char 64->64 convolutions and the three
FC GEMMs (and any conv with Cin*kh*kw >= 64 and Cout >= 16); HBM for everything else.
"""
from .nn import layers as L


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def layer_work(layer, in_shape, direction='forward'):
    """-> {'bound': 'hbm'|'tensor', 'bytes': algorithmic bytes, 'flops': algorithmic flops}."""
    mult = 1 if direction == 'forward' else 2          # backward = dgrad + wgrad
    if isinstance(layer, L.Convolutional2D):
        n, h, w, cin = in_shape
        _, ho, wo, cout = layer.get_output_shapes(in_shape)[0]
        kh, kw = layer.kernel_size
        flops = 2 * n * ho * wo * kh * kw * cin * cout * mult
        nbytes = 4 * (n * h * w * cin + n * ho * wo * cout + kh * kw * cin * cout) * mult
        tensor = cin * kh * kw >= 64 and cout >= 16
        return {'bound': 'tensor' if tensor else 'hbm', 'bytes': nbytes, 'flops': flops}
    if isinstance(layer, L.FullyConnected):
        b = in_shape[0]
        flops = 2 * b * (layer.n_input + 1) * layer.n_output * mult
        nbytes = 4 * (b * layer.n_input + b * layer.n_output + (layer.n_input + 1) * layer.n_output) * mult
        return {'bound': 'tensor', 'bytes': nbytes, 'flops': flops}
    numel = _numel(in_shape)
    if isinstance(layer, L.Upsample2D):
        sy, sx = layer.scale_factor
        return {'bound': 'hbm', 'bytes': 4 * numel * (1 + sy * sx), 'flops': 0}
    if isinstance(layer, L.Conv2DToBatchedFixedWidthed):
        return {'bound': 'hbm', 'bytes': 4 * numel * (1 + layer.width), 'flops': 0}
    if isinstance(layer, L.MaxPool2D):
        out = _numel(layer.get_output_shapes(in_shape)[0])
        kh, kw = layer.kernel_size
        return {'bound': 'hbm', 'bytes': 4 * (numel + out) + out * kh * kw, 'flops': 0}
    if isinstance(layer, (L.LeakyRelu, L.Sigmoid)):
        return {'bound': 'hbm', 'bytes': (8 if direction == 'forward' else 12) * numel, 'flops': numel}
    return {'bound': 'hbm', 'bytes': 0, 'flops': 0}         # Flatten / Noop: views


def model_input_shapes(model, input_shape):
    """{leaf layer name: its (first) input shape} for a model fed `input_shape`."""
    _, all_shapes = model.get_all_output_shapes([input_shape])
    shapes = {}
    for name in model._order:
        src = model.relations[name][0]
        shapes[name] = tuple(input_shape) if isinstance(src, int) else tuple(all_shapes[src][0])
    return shapes


def plan_work(model, input_shape, training=False):
    """{tracked layer name: work} for the model's execution plan: a fused step is one kernel and
    is tracked under its LAST layer's name; its algorithmic bytes are the step's external input
    + output + weights (the fused-away intermediates cost nothing), its flops the sum."""
    shapes = model_input_shapes(model, input_shape)
    _, all_shapes = model.get_all_output_shapes([input_shape])
    out = {}
    hook = getattr(model, 'infer_fusion', None)
    if (not training and hook is not None and getattr(model, '_fusion_planned', False)
            and hook._blocks(model) is not None and input_shape[1] % 4 == 0 and input_shape[2] % 4 == 0):
        # whole-network kernel (HourglassFusion): one launch, tracked under the last layer's name
        names = list(hook.chain)
        flops = sum(layer_work(model.layers[n], shapes[n], 'forward')['flops'] for n in names
                    if isinstance(model.layers[n], L.Convolutional2D))
        wbytes = sum(4 * model.layers[n].count_parameters() for n in names
                     if isinstance(model.layers[n], L.Convolutional2D))
        nbytes = 4 * (_numel(input_shape) + _numel(all_shapes[names[-1]][0])) + wbytes
        return {names[-1]: {'bound': 'hbm', 'bytes': nbytes, 'flops': flops, 'fused': names}}
    for step in (model._plan_train if training else model._plan_infer):
        names = [n for n in step[1:] if n is not None]
        if step[0] == 'layer':
            out[names[0]] = layer_work(model.layers[names[0]], shapes[names[0]], 'forward')
            continue
        flops, wbytes, bound = 0, 0, 'hbm'
        for n in names:
            layer = model.layers[n]
            wk = layer_work(layer, shapes[n], 'forward')
            if isinstance(layer, (L.Convolutional2D, L.FullyConnected)):
                flops += wk['flops']
                wbytes += 4 * layer.count_parameters()
                if wk['bound'] == 'tensor':
                    bound = 'tensor'
        nbytes = 4 * (_numel(shapes[names[0]]) + _numel(all_shapes[names[-1]][0])) + wbytes
        out[names[-1]] = {'bound': bound, 'bytes': nbytes, 'flops': flops, 'fused': names}
    return out
