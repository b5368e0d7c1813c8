# compute-sanitizer racecheck + memcheck over the experimental split engine (front + tail kernels)
mkdir -p gpurun_out
cat > /tmp/san_split.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from deepbinner_b200.model import B200Model
z = np.load('tests/golden/fixture_reads.npz')
sigs = [z['signal_%d' % i] for i in range(7)]
m = B200Model('deepbinner_b200/models/EXP-NBD103_read_starts.dbnw')
m.set_engine('tcgen05-split')
calls, probs = m.call_batch(sigs[:3], 'end', 1024, 0.5)
p = m.predict(np.random.RandomState(0).randn(9, 1024).astype(np.float32))
print('split', calls.tolist(), float(p.sum()))
PY
for tool in racecheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_split.py > gpurun_out/r01_sanitizer_split_$tool.txt 2>&1
  tail -3 gpurun_out/r01_sanitizer_split_$tool.txt
done
