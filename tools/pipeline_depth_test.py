import sys, time, types, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from deepbinner_b200 import classify as cls
from deepbinner_b200.model import B200Model
start = B200Model(bench.model_path()); end = B200Model(str(bench.ROOT/'deepbinner_b200/models/EXP-NBD103_read_ends.dbnw'))
ns = types.SimpleNamespace(scan_size=6144.0, batch_size=256, score_diff=0.5, require_either=True, require_start=False, require_both=False, verbose=False)
ids=['r%d'%i for i in range(256)]
b=[(ids, bench.ragged_reads(256, 7+k)) for k in range(4)]
for rep in range(2):
  for depth in (1,2,3):
    cls.PIPELINE_DEPTH = depth
    r = bench.pipeline_rate(cls, b, start, end, ns, 13, reps=10)
    print('depth', depth, 'config2 windows/s %.0f' % (24*r), flush=True)
