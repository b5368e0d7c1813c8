"""
Host-side signal normalisation, same contract as reference `trim_signal.py:61-69 normalise`:
z-score with the population standard deviation; an empty signal is returned unchanged; a constant
signal only has its mean removed.  Only used when call_batch drives a foreign `.predict` model
(seam b1); the fused GPU path does this on the device.
"""

import numpy as np


def normalise(signal):
    signal = np.asarray(signal)
    if signal.size == 0:
        return signal
    centred = signal - signal.mean()
    spread = signal.std()
    return centred / spread if spread > 0.0 else centred
