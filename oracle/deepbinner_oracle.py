"""
ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (numpy, float64 by default) of the Deepbinner barcode-classification hot path:
raw signal -> windows -> z-score -> 1-D CNN -> softmax -> per-read merge -> call.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
anything under `oracle/`; the product (`deepbinner_b200/`) never does and fails loudly when its
CUDA library is missing.

The arithmetic of the reference lives in un-vendored third-party packages (`tensorflow` 1.x and
`keras` 2.1.4 per the `keras_version`/`backend` attributes of the shipped model files;
`requirements.txt:3,7`, call sites `classify.py:90` and `classify.py:361`), neither of which is
installed or installable here.  This file therefore restates the published Keras/TensorFlow
inference semantics for exactly the graph that `network_architecture.py:18-95` builds, and the
reference's own Python for everything around it:

  normalise                     <- trim_signal.py:61-69
  make_windows                  <- classify.py:337-358   (windowing / padding inside call_batch)
  forward (CNN)                 <- network_architecture.py:18-95 + Keras/TF layer semantics
  merge_steps                   <- classify.py:363-377
  make_sum_to_one               <- classify.py:387-393
  get_barcode_call_from_probabilities <- classify.py:285-295
  combine_calls                 <- classify.py:298-322
  call_batch                    <- classify.py:325-384

PARITY PINNING: pinned against every golden the reference's tests hold for this path - the 28 calls
of tests/test_classify.py:104-180, the 2-d.p. probability rows at :213-217/:249-253/:287-296, the
truth table of tests/test_combine_calls.py:27-51, and 107,197 parameters
(tests/test_network_architecture.py:37); plus, in the authoring container, the reference's own
`call_batch`/`combine_calls` imported from /root/reference (h5py/keras/tensorflow stubbed) driven
by this forward pass (see tests/test_oracle_pinning.py and tests/golden/make_golden.py).  The
reference pins no probability below 2 decimals, so the 1e-3 comparison between the CUDA path and
this oracle rests on this file being a faithful restatement of Keras semantics.
"""

import struct

import numpy as np

BN_EPSILON = 1e-3  # Keras BatchNormalization epsilon recorded in the model_config of models/*

_MAGIC = b'DBNWGT1\x00'


# --------------------------------------------------------------------------------------------
# Weights (own reader for the DBNW blob so the oracle does not depend on product code)
# --------------------------------------------------------------------------------------------
def load_weights(path, dtype=np.float64):
    """Read a DBNW weight blob -> dict with 'input_size', 'n_classes' and '<layer>/<tensor>'."""
    with open(str(path), 'rb') as fh:
        blob = fh.read()
    magic, version, input_size, n_classes, n_tensors = struct.unpack_from('<8sIIII', blob, 0)
    assert magic == _MAGIC and version == 1, 'not a DBNW v1 blob'
    entry = struct.Struct('<48sI3IQQ')
    data_start = 24 + n_tensors * entry.size
    out = {'input_size': input_size, 'n_classes': n_classes}
    for i in range(n_tensors):
        name, ndim, d0, d1, d2, off, count = entry.unpack_from(blob, 24 + i * entry.size)
        arr = np.frombuffer(blob, dtype='<f4', count=count, offset=data_start + 4 * off)
        out[name.split(b'\x00')[0].decode()] = arr.reshape((d0, d1, d2)[:ndim]).astype(dtype)
    return out


# --------------------------------------------------------------------------------------------
# Signal handling
# --------------------------------------------------------------------------------------------
def normalise(signal):
    """trim_signal.py:61-69 - z-score with population stdev; empty -> unchanged; sd 0 -> x-mean."""
    if len(signal) == 0:
        return signal
    mean = np.mean(signal)
    stdev = np.std(signal)
    if stdev > 0.0:
        return (signal - mean) / stdev
    else:
        return signal - mean


def make_windows(signals, input_size, step_index, side):
    """classify.py:337-358 - the [N, input_size] float64 network input for scan step `step_index`."""
    step_size = input_size // 2
    sig_start = step_index * step_size
    sig_end = sig_start + input_size
    out = np.zeros((len(signals), input_size), dtype=np.float64)
    for i, signal in enumerate(signals):
        if side == 'start':
            piece = signal[sig_start:sig_end]
        else:
            assert side == 'end'
            a = max(len(signal) - sig_end, 0)
            b = max(len(signal) - sig_start, 0)
            piece = signal[a:b]
        piece = normalise(np.asarray(piece))
        n = len(piece)
        if n == 0:
            continue
        if side == 'start':
            out[i, :n] = piece            # zero pad on the right  (classify.py:354-355)
        else:
            out[i, input_size - n:] = piece   # zero pad on the left (classify.py:356-357)
    return out


# --------------------------------------------------------------------------------------------
# Keras / TensorFlow layer semantics
# --------------------------------------------------------------------------------------------
def conv1d_relu(x, kernel, bias, stride=1):
    """Keras Conv1D(padding='same', activation='relu') on channels-last x[N, L, Cin].

    TF SAME padding: Lout = ceil(L/s); pad_total = max((Lout-1)*s + k - L, 0); left = total // 2
    (so k=3,s=2 on even L pads 0 left / 1 right).  Cross-correlation, kernel [k, Cin, Cout]."""
    n, length, cin = x.shape
    k, kcin, cout = kernel.shape
    assert kcin == cin
    lout = -(-length // stride)
    pad_total = max((lout - 1) * stride + k - length, 0)
    pad_left = pad_total // 2
    pad_right = pad_total - pad_left
    xp = np.pad(x, ((0, 0), (pad_left, pad_right), (0, 0)))
    cols = np.empty((n, lout, k * cin), dtype=x.dtype)
    for t in range(k):
        cols[:, :, t * cin:(t + 1) * cin] = xp[:, t:t + (lout - 1) * stride + 1:stride, :]
    y = cols.reshape(n * lout, k * cin) @ kernel.reshape(k * cin, cout)
    y = y.reshape(n, lout, cout) + bias
    return np.maximum(y, 0.0)


def batch_norm(x, w, name):
    """Keras BatchNormalization at inference: gamma*(x-mean)/sqrt(var+eps)+beta, per channel."""
    gamma, beta = w[name + '/gamma'], w[name + '/beta']
    mean, var = w[name + '/moving_mean'], w[name + '/moving_variance']
    return gamma * (x - mean) / np.sqrt(var + x.dtype.type(BN_EPSILON)) + beta


def max_pool2(x):
    """MaxPooling1D(pool_size=2): stride 2, 'valid'."""
    n, length, c = x.shape
    return x[:, :length // 2 * 2, :].reshape(n, length // 2, 2, c).max(axis=2)


def avg_pool3_same(x):
    """AveragePooling1D(3, strides=1, padding='same'): TF divides by the number of in-range taps."""
    n, length, c = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (0, 0)))
    s = xp[:, 0:length] + xp[:, 1:length + 1] + xp[:, 2:length + 2]
    cnt = np.full((1, length, 1), 3.0, dtype=x.dtype)
    cnt[0, 0, 0] = 2.0
    cnt[0, -1, 0] = 2.0
    if length == 1:
        cnt[0, 0, 0] = 1.0
    return s / cnt


def softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def forward(w, x, return_logits=False, taps=None):
    """The network of network_architecture.py:18-95 at inference (GaussianNoise and Dropout are
    identity).  x: [N, input_size] or [N, input_size, 1]; returns softmax rows [N, n_classes].
    If `taps` is a dict it receives the intermediate tensors (for per-layer kernel tests)."""
    dtype = w['conv1d_1/kernel'].dtype
    if taps is None:
        taps = {}
    x = np.asarray(x, dtype=dtype)
    if x.ndim == 2:
        x = x[:, :, None]

    def conv(name, t, stride=1):
        return conv1d_relu(t, w[name + '/kernel'], w[name + '/bias'], stride)

    x = conv('conv1d_1', x, stride=2)                       # :26
    x = batch_norm(x, w, 'batch_normalization_1')           # :28
    taps['bn1'] = x
    x = conv('conv1d_2', x)                                 # :32-34
    taps['conv2'] = x
    x = conv('conv1d_3', x)
    taps['conv3'] = x
    x = conv('conv1d_4', x)
    x = max_pool2(x)                                        # :35
    x = batch_norm(x, w, 'batch_normalization_2')           # :37
    taps['bn2'] = x
    x = conv('conv1d_5', x)                                 # :41 bottleneck (k=1)
    taps['conv5'] = x
    x = conv('conv1d_6', x)                                 # :44-45
    taps['conv6'] = x
    x = conv('conv1d_7', x)
    x = max_pool2(x)                                        # :46
    x = batch_norm(x, w, 'batch_normalization_3')           # :48
    taps['bn3'] = x
    x = conv('conv1d_8', x)                                 # :52-53
    taps['conv8'] = x
    x = conv('conv1d_9', x)
    x = max_pool2(x)                                        # :54
    x = batch_norm(x, w, 'batch_normalization_4')           # :56
    taps['bn4'] = x
    taps['avgpool'] = avg_pool3_same(x)
    taps['conv12'] = conv('conv1d_12', x)
    taps['conv14'] = conv('conv1d_14', x)
    taps['conv15'] = conv('conv1d_15', taps['conv14'])
    x1 = conv('conv1d_10', avg_pool3_same(x))               # :60-61
    x2 = conv('conv1d_11', x)                               # :62
    x3 = conv('conv1d_13', conv('conv1d_12', x))            # :63-64
    x4 = conv('conv1d_16', conv('conv1d_15', conv('conv1d_14', x)))   # :65-67
    x = np.concatenate([x1, x2, x3, x4], axis=2)            # :68
    x = max_pool2(x)                                        # :69
    x = batch_norm(x, w, 'batch_normalization_5')           # :71
    taps['bn5'] = x
    x = conv('conv1d_17', x, stride=2)                      # :75
    x = batch_norm(x, w, 'batch_normalization_6')           # :77
    taps['bn6'] = x
    x = conv('conv1d_18', x)                                # :81-82
    taps['conv18'] = x
    x = conv('conv1d_19', x)
    x = max_pool2(x)                                        # :83
    x = batch_norm(x, w, 'batch_normalization_7')           # :85
    taps['bn7'] = x
    x = conv('conv1d_20', x)                                # :89 (ReLU before the pooling)
    logits = x.mean(axis=1)                                 # :90 GlobalAveragePooling1D
    if return_logits:
        return logits
    return softmax(logits)                                  # :91


class OracleModel:
    """Object with the Keras surface `call_batch` uses (classify.py:92-99, :361): `.inputs`,
    `.outputs`, `.predict(x, batch_size)` returning fresh float32 rows."""

    class _T:
        def __init__(self, shape):
            self.shape = shape

    def __init__(self, weights_path, dtype=np.float64):
        self.w = load_weights(weights_path, dtype)
        self.inputs = [self._T((None, self.w['input_size'], 1))]
        self.outputs = [self._T((None, self.w['n_classes']))]

    def predict(self, x, batch_size=256):
        x = np.asarray(x)
        out = np.empty((x.shape[0], self.w['n_classes']), dtype=np.float32)
        for s in range(0, x.shape[0], batch_size):
            out[s:s + batch_size] = forward(self.w, x[s:s + batch_size])
        return out


# --------------------------------------------------------------------------------------------
# Per-read merge and call logic
# --------------------------------------------------------------------------------------------
def merge_steps(step_probs):
    """classify.py:363-377 - step_probs [steps, N, C] -> [N, C]: class 0 = min over steps,
    classes >= 1 = max over steps (the override at :376-377 can never fire after the merge)."""
    step_probs = np.asarray(step_probs)
    merged = step_probs.max(axis=0)
    merged[:, 0] = step_probs[:, :, 0].min(axis=0)
    return merged


def make_sum_to_one(probabilities):
    """classify.py:387-393."""
    no_barcode_prob = probabilities[0]
    all_barcode_probs = 1.0 - no_barcode_prob
    factor = all_barcode_probs / sum(probabilities[1:])
    probabilities = [p * factor for p in probabilities]
    probabilities[0] = no_barcode_prob
    return probabilities


def get_barcode_call_from_probabilities(probabilities, score_diff_threshold):
    """classify.py:285-295 (stable sort => ties go to the lower class index)."""
    ranked = sorted(enumerate(probabilities), key=lambda x: x[1], reverse=True)
    best, second_best = ranked[0], ranked[1]
    if best[0] == 0:
        return 'none'
    if best[1] - second_best[1] >= score_diff_threshold:
        return str(best[0])
    return 'none'


def combine_calls(start_call, end_call, require_either=False, require_start=False,
                  require_both=False):
    """classify.py:298-322."""
    if require_both:
        return start_call if start_call == end_call else 'none'
    if require_start:
        if start_call == end_call:
            return start_call
        if start_call == 'none':
            return 'none'
        if end_call == 'none':
            return start_call
        return 'none'
    assert require_either
    if start_call == end_call:
        return start_call
    if start_call == 'none':
        return end_call
    if end_call == 'none':
        return start_call
    return 'none'


def call_batch(w, signals, side, scan_size=6144, score_diff=0.5, return_steps=False):
    """classify.py:325-384 - returns (calls, probabilities[N][C]) (+ per-step softmax rows)."""
    input_size = w['input_size']
    step_size = input_size // 2
    steps = int(scan_size / step_size)
    assert steps * step_size == scan_size
    step_probs = []
    for s in range(steps):
        x = make_windows(signals, input_size, s, side)
        # Keras casts the float64 input to float32 (floatx) before the graph (Appendix B.7)
        step_probs.append(forward(w, x.astype(np.float32)).astype(np.float32))
    merged = merge_steps(step_probs)
    calls, probs = [], []
    for row in merged:
        p = make_sum_to_one(list(row))
        probs.append(p)
        calls.append(get_barcode_call_from_probabilities(p, score_diff))
    if return_steps:
        return calls, probs, np.asarray(step_probs)
    return calls, probs
