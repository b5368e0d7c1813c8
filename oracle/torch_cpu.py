"""
ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product path.

PyTorch-CPU float32 restatement of the same graph as `oracle/deepbinner_oracle.py:forward`
(reference `network_architecture.py:18-95`, executed by `model.predict` at `classify.py:361`).
It exists for one reason: the reference's own TensorFlow-CPU `model.predict` cannot be run in this
image (tensorflow/keras/h5py absent, no network), so this multi-threaded fp32 CPU forward pass is
the stand-in CPU baseline that `bench.py` times on the GPU box's host cores (`cpu_baseline`,
`--impl reference`).  It is validated against the numpy fp64 oracle in tests/test_oracle_pinning.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.
"""

import numpy as np
import torch
import torch.nn.functional as F

from . import deepbinner_oracle as orc


class TorchCpuModel:
    """fp32 channels-first CPU forward pass; `.predict(x[N,1024(,1)]) -> float32 [N, C]`."""

    def __init__(self, weights_path, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        w = orc.load_weights(weights_path, np.float32)
        self.input_size = w['input_size']
        self.n_classes = w['n_classes']
        self.t = {}
        for name, arr in w.items():
            if not isinstance(arr, np.ndarray):
                continue
            t = torch.from_numpy(np.ascontiguousarray(arr))
            if name.endswith('/kernel'):
                t = t.permute(2, 1, 0).contiguous()   # [k,Cin,Cout] -> [Cout,Cin,k]
            self.t[name] = t
        self.bn = {}
        for i in range(1, 8):
            n = 'batch_normalization_{}'.format(i)
            scale = self.t[n + '/gamma'] / torch.sqrt(self.t[n + '/moving_variance'] + orc.BN_EPSILON)
            shift = self.t[n + '/beta'] - self.t[n + '/moving_mean'] * scale
            self.bn[i] = (scale.view(1, -1, 1), shift.view(1, -1, 1))

    def _conv(self, name, x, stride=1):
        k = self.t[name + '/kernel']
        b = self.t[name + '/bias']
        ksize = k.shape[2]
        if ksize == 3 and stride == 1:
            x = F.pad(x, (1, 1))
        elif ksize == 3 and stride == 2:
            x = F.pad(x, (0, 1))          # TF SAME, even length: pad right only
        return F.relu(F.conv1d(x, k, b, stride=stride))

    def _bn(self, i, x):
        scale, shift = self.bn[i]
        return x * scale + shift

    @torch.no_grad()
    def forward(self, x):
        c = self._conv
        x = self._bn(1, c('conv1d_1', x, 2))
        x = c('conv1d_4', c('conv1d_3', c('conv1d_2', x)))
        x = self._bn(2, F.max_pool1d(x, 2))
        x = c('conv1d_7', c('conv1d_6', c('conv1d_5', x)))
        x = self._bn(3, F.max_pool1d(x, 2))
        x = c('conv1d_9', c('conv1d_8', x))
        x = self._bn(4, F.max_pool1d(x, 2))
        x1 = c('conv1d_10', F.avg_pool1d(x, 3, stride=1, padding=1, count_include_pad=False))
        x2 = c('conv1d_11', x)
        x3 = c('conv1d_13', c('conv1d_12', x))
        x4 = c('conv1d_16', c('conv1d_15', c('conv1d_14', x)))
        x = torch.cat([x1, x2, x3, x4], dim=1)
        x = self._bn(5, F.max_pool1d(x, 2))
        x = self._bn(6, c('conv1d_17', x, 2))
        x = c('conv1d_19', c('conv1d_18', x))
        x = self._bn(7, F.max_pool1d(x, 2))
        x = c('conv1d_20', x)
        return torch.softmax(x.mean(dim=2), dim=1)

    def predict(self, x, batch_size=256):
        x = np.asarray(x, dtype=np.float32).reshape(len(x), 1, self.input_size)
        out = np.empty((len(x), self.n_classes), dtype=np.float32)
        for s in range(0, len(x), batch_size):
            out[s:s + batch_size] = self.forward(torch.from_numpy(x[s:s + batch_size])).numpy()
        return out
