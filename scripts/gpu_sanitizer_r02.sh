# compute-sanitizer passes over a small run of both engines (memcheck + racecheck + synccheck)
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from deepbinner_b200.model import B200Model
z = np.load('tests/golden/fixture_reads.npz')
sigs = [z['signal_%d' % i] for i in range(7)]
m = B200Model('deepbinner_b200/models/EXP-NBD103_read_starts.dbnw')
for eng in ('tcgen05', 'fp32'):
    m.set_engine(eng)
    calls, probs = m.call_batch(sigs[:3], 'end', 1024, 0.5)
    calls2, probs2 = m.call_batch(sigs, 'start', 6144, 0.5)
    p = m.predict(np.random.RandomState(0).randn(5, 1024).astype(np.float32))
    big = m.predict(np.random.RandomState(1).randn(420, 1024).astype(np.float32))   # 210 window pairs: some CTAs of the persistent kernel run two
    calls3, probs3 = m.call_batch(sigs * 6, 'start', 6144, 0.5)                       # 252 window pairs through the fused form
    print(eng, calls.tolist(), float(p.sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
