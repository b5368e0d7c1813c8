"""
Model weights: Keras HDF5 model file -> flat weight blob ("DBNW" format) consumed by the C-ABI.

Replaces `keras.models.load_model` at reference `classify.py:86-103 load_trained_model`: the three
shipped model files (`models/EXP-NBD103_read_starts`, `..._read_ends`, `SQK-RBK004_read_starts`)
are Keras 2.1.4 HDF5 files whose `model_config` attribute describes the graph of
`network_architecture.py:18-95 build_network` and whose `model_weights/<layer>/<layer>/<name>:0`
datasets hold the fp32 parameters.  We verify the graph is that topology and pack the parameters
into one contiguous little-endian blob:

    header   : magic 'DBNWGT1\\0' | u32 version=1 | u32 input_size | u32 n_classes | u32 n_tensors
    table    : n_tensors x { char name[48] | u32 ndim | u32 dims[3] | u64 offset (floats) | u64 count }
    payload  : float32 values, tensors in table order

Tensor names are '<layer>/<kernel|bias|gamma|beta|moving_mean|moving_variance>'.  Conv kernels keep
the Keras layout [k, Cin, Cout].
"""

import json
import struct

import numpy as np

from . import hdf5_lite

MAGIC = b'DBNWGT1\x00'
VERSION = 1
_HEADER = struct.Struct('<8sIIII')
_ENTRY = struct.Struct('<48sI3IQQ')

# (name, kernel_size, stride, Cin, Cout) for the 20 Conv1D layers of build_network
# (reference network_architecture.py:22-95).  Cout of conv1d_20 = class count.
CONV_SPECS = [
    ('conv1d_1', 3, 2, 1, 48), ('conv1d_2', 3, 1, 48, 48), ('conv1d_3', 3, 1, 48, 48),
    ('conv1d_4', 3, 1, 48, 48), ('conv1d_5', 1, 1, 48, 16), ('conv1d_6', 3, 1, 16, 48),
    ('conv1d_7', 3, 1, 48, 48), ('conv1d_8', 3, 1, 48, 48), ('conv1d_9', 3, 1, 48, 48),
    ('conv1d_10', 1, 1, 48, 48), ('conv1d_11', 1, 1, 48, 48), ('conv1d_12', 1, 1, 48, 16),
    ('conv1d_13', 3, 1, 16, 48), ('conv1d_14', 1, 1, 48, 16), ('conv1d_15', 3, 1, 16, 48),
    ('conv1d_16', 3, 1, 48, 48), ('conv1d_17', 3, 2, 192, 48), ('conv1d_18', 3, 1, 48, 48),
    ('conv1d_19', 3, 1, 48, 48), ('conv1d_20', 1, 1, 48, None),
]
BN_CHANNELS = [48, 48, 48, 48, 192, 48, 48]
BN_EPSILON = 1e-3

# Inbound wiring of the saved graph (layer -> its input layer), used to validate model_config.
_EXPECTED_INBOUND = {
    'gaussian_noise_1': ['input_1'], 'conv1d_1': ['gaussian_noise_1'],
    'batch_normalization_1': ['conv1d_1'], 'dropout_1': ['batch_normalization_1'],
    'conv1d_2': ['dropout_1'], 'conv1d_3': ['conv1d_2'], 'conv1d_4': ['conv1d_3'],
    'max_pooling1d_1': ['conv1d_4'], 'batch_normalization_2': ['max_pooling1d_1'],
    'dropout_2': ['batch_normalization_2'], 'conv1d_5': ['dropout_2'], 'conv1d_6': ['conv1d_5'],
    'conv1d_7': ['conv1d_6'], 'max_pooling1d_2': ['conv1d_7'],
    'batch_normalization_3': ['max_pooling1d_2'], 'dropout_3': ['batch_normalization_3'],
    'conv1d_8': ['dropout_3'], 'conv1d_9': ['conv1d_8'], 'max_pooling1d_3': ['conv1d_9'],
    'batch_normalization_4': ['max_pooling1d_3'], 'dropout_4': ['batch_normalization_4'],
    'average_pooling1d_1': ['dropout_4'], 'conv1d_10': ['average_pooling1d_1'],
    'conv1d_11': ['dropout_4'], 'conv1d_12': ['dropout_4'], 'conv1d_13': ['conv1d_12'],
    'conv1d_14': ['dropout_4'], 'conv1d_15': ['conv1d_14'], 'conv1d_16': ['conv1d_15'],
    'concatenate_1': ['conv1d_10', 'conv1d_11', 'conv1d_13', 'conv1d_16'],
    'max_pooling1d_4': ['concatenate_1'], 'batch_normalization_5': ['max_pooling1d_4'],
    'dropout_5': ['batch_normalization_5'], 'conv1d_17': ['dropout_5'],
    'batch_normalization_6': ['conv1d_17'], 'dropout_6': ['batch_normalization_6'],
    'conv1d_18': ['dropout_6'], 'conv1d_19': ['conv1d_18'], 'max_pooling1d_5': ['conv1d_19'],
    'batch_normalization_7': ['max_pooling1d_5'], 'dropout_7': ['batch_normalization_7'],
    'conv1d_20': ['dropout_7'], 'global_average_pooling1d_1': ['conv1d_20'],
    'softmax_1': ['global_average_pooling1d_1'],
}


class ModelFormatError(ValueError):
    pass


def _relative_names(layers):
    """Keras numbers layers globally per process (conv1d_21.. for a second model).  Map each layer
    name to its name relative to the first occurrence of its kind (conv1d_1, ...)."""
    counters = {}
    for layer in layers:
        name = layer['name']
        stem, _, num = name.rpartition('_')
        counters.setdefault(stem, []).append(int(num))
    mapping = {}
    for stem, nums in counters.items():
        base = min(nums)
        for n in nums:
            mapping['{}_{}'.format(stem, n)] = '{}_{}'.format(stem, n - base + 1)
    return mapping


def check_topology(model_config):
    """Validate that a Keras model_config JSON is the Deepbinner network; return
    (input_size, n_classes, name_map) where name_map maps saved layer names -> canonical names."""
    cfg = model_config['config']
    layers = cfg['layers']
    name_map = _relative_names(layers)
    by_name = {name_map[l['name']]: l for l in layers}
    if len(layers) != 45:
        raise ModelFormatError('expected the 45-layer Deepbinner network, found {} layers'
                               .format(len(layers)))
    for name, inbound in _EXPECTED_INBOUND.items():
        if name not in by_name:
            raise ModelFormatError('layer {} missing from model'.format(name))
        got = [name_map[x[0]] for x in by_name[name]['inbound_nodes'][0]]
        if got != inbound:
            raise ModelFormatError('layer {} is wired to {}, expected {}'.format(name, got, inbound))
    in_shape = by_name['input_1']['config']['batch_input_shape']
    input_size = int(in_shape[1])
    if int(in_shape[2]) != 1:
        raise ModelFormatError('model input must have one channel')
    n_classes = None
    for name, k, s, cin, cout in CONV_SPECS:
        c = by_name[name]['config']
        if cout is None:
            n_classes = cout = int(c['filters'])
        pad_ok = c['padding'] == 'same' or k == 1
        if (int(c['filters']), int(c['kernel_size'][0]), int(c['strides'][0])) != (cout, k, s) \
                or c['activation'] != 'relu' or not c['use_bias'] or not pad_ok \
                or int(c['dilation_rate'][0]) != 1:
            raise ModelFormatError('layer {} has unexpected hyper-parameters'.format(name))
    for i in range(1, 8):
        c = by_name['batch_normalization_{}'.format(i)]['config']
        if abs(float(c['epsilon']) - BN_EPSILON) > 1e-12 or not c['center'] or not c['scale'] \
                or int(c['axis']) not in (-1, 2):
            raise ModelFormatError('batch_normalization_{} has unexpected config'.format(i))
    for i in range(1, 6):
        c = by_name['max_pooling1d_{}'.format(i)]['config']
        if int(c['pool_size'][0]) != 2 or int(c['strides'][0]) != 2:
            raise ModelFormatError('max_pooling1d_{} has unexpected config'.format(i))
    c = by_name['average_pooling1d_1']['config']
    if int(c['pool_size'][0]) != 3 or int(c['strides'][0]) != 1 or c['padding'] != 'same':
        raise ModelFormatError('average_pooling1d_1 has unexpected config')
    return input_size, n_classes, name_map


def tensors_from_keras_file(path):
    """Parse a Keras HDF5 model file -> (input_size, n_classes, ordered dict name -> fp32 array)."""
    try:
        f = hdf5_lite.open_file(path)
    except hdf5_lite.Hdf5Error as e:
        raise ModelFormatError(str(e))
    with f:
        if 'model_config' not in f.attrs or 'model_weights' not in f.keys():
            raise ModelFormatError('{} is not a Keras model file'.format(path))
        config = json.loads(f.attrs['model_config'].decode('utf-8'))
        input_size, n_classes, name_map = check_topology(config)
        inverse = {v: k for k, v in name_map.items()}
        mw = f['model_weights']
        tensors = {}

        def fetch(layer, weight):
            saved = inverse[layer]
            arr = mw['{0}/{0}/{1}:0'.format(saved, weight)].read()
            return np.ascontiguousarray(arr, dtype='<f4')

        for name, k, s, cin, cout in CONV_SPECS:
            cout = n_classes if cout is None else cout
            kern = fetch(name, 'kernel')
            bias = fetch(name, 'bias')
            if kern.shape != (k, cin, cout) or bias.shape != (cout,):
                raise ModelFormatError('{}: unexpected weight shapes {} {}'
                                       .format(name, kern.shape, bias.shape))
            tensors[name + '/kernel'] = kern
            tensors[name + '/bias'] = bias
        for i, ch in enumerate(BN_CHANNELS, start=1):
            name = 'batch_normalization_{}'.format(i)
            for w in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                arr = fetch(name, w)
                if arr.shape != (ch,):
                    raise ModelFormatError('{}: unexpected shape {}'.format(name, arr.shape))
                tensors['{}/{}'.format(name, w)] = arr
    return input_size, n_classes, tensors


def pack_blob(input_size, n_classes, tensors):
    names = list(tensors.keys())
    table = b''
    payload = []
    offset = 0
    for name in names:
        arr = np.ascontiguousarray(tensors[name], dtype='<f4')
        dims = list(arr.shape) + [0] * (3 - arr.ndim)
        table += _ENTRY.pack(name.encode('ascii'), arr.ndim, dims[0], dims[1], dims[2],
                             offset, arr.size)
        payload.append(arr.tobytes())
        offset += arr.size
    header = _HEADER.pack(MAGIC, VERSION, input_size, n_classes, len(names))
    return header + table + b''.join(payload)


def unpack_blob(blob):
    """Inverse of pack_blob -> (input_size, n_classes, dict name -> fp32 array)."""
    blob = bytes(blob)
    if len(blob) < _HEADER.size or blob[:8] != MAGIC:
        raise ModelFormatError('not a DBNW weight blob')
    _, version, input_size, n_classes, n_tensors = _HEADER.unpack_from(blob, 0)
    if version != VERSION:
        raise ModelFormatError('unsupported DBNW version {}'.format(version))
    data_start = _HEADER.size + n_tensors * _ENTRY.size
    tensors = {}
    for i in range(n_tensors):
        name, ndim, d0, d1, d2, offset, count = _ENTRY.unpack_from(blob, _HEADER.size + i * _ENTRY.size)
        name = name.split(b'\x00')[0].decode('ascii')
        shape = (d0, d1, d2)[:ndim]
        arr = np.frombuffer(blob, dtype='<f4', count=count, offset=data_start + 4 * offset)
        tensors[name] = arr.reshape(shape)
    return input_size, n_classes, tensors


def is_blob_file(path):
    try:
        with open(str(path), 'rb') as fh:
            return fh.read(8) == MAGIC
    except (IOError, OSError):
        return False


def load_blob(path):
    """Load a model file (DBNW blob or Keras HDF5) and return the DBNW blob bytes."""
    if is_blob_file(path):
        with open(str(path), 'rb') as fh:
            blob = fh.read()
        unpack_blob(blob)
        return blob
    input_size, n_classes, tensors = tensors_from_keras_file(path)
    return pack_blob(input_size, n_classes, tensors)


def parameter_count(blob):
    _, _, tensors = unpack_blob(blob)
    return int(sum(t.size for t in tensors.values()))
