"""
deepbinner_b200: B200-native (sm_100a CUDA) drop-in for the classification hot path of
rrwick/Deepbinner - `classify.py:call_batch` / `model.predict` / `classify_fast5_files`.

The product path always runs on the GPU through libdeepbinner_b200.so (C ABI in
include/deepbinner_b200.h); there is no CPU fallback.
"""

__version__ = '0.1.0'
