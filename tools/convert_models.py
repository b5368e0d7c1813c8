#!/usr/bin/env python3
"""
Convert the reference's Keras HDF5 model files into DBNW weight blobs under
deepbinner_b200/models/ (the trained weights are the one thing reused from the reference; see
reference models/README.md).  Usage: python tools/convert_models.py [/root/reference/models]
"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from deepbinner_b200 import weights  # noqa: E402


def main():
    src = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else '/root/reference/models')
    dst = ROOT / 'deepbinner_b200' / 'models'
    dst.mkdir(exist_ok=True)
    for name in ('EXP-NBD103_read_starts', 'EXP-NBD103_read_ends', 'SQK-RBK004_read_starts'):
        blob = weights.load_blob(src / name)
        input_size, n_classes, _ = weights.unpack_blob(blob)
        out = dst / (name + '.dbnw')
        out.write_bytes(blob)
        print('{}: input_size={} classes={} params={} -> {} ({} bytes)'.format(
            name, input_size, n_classes, weights.parameter_count(blob), out, len(blob)))


if __name__ == '__main__':
    main()
