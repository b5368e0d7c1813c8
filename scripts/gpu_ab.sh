# A/B of kernel variants: for the default library and every variants/*.so: quick parity subset, timeline total, short bench.
mkdir -p gpurun_out
for lib in deepbinner_b200/libdeepbinner_b200.so $(ls variants/*.so 2>/dev/null); do
  name=$(basename $lib .so)
  export DEEPBINNER_B200_LIB=$PWD/$lib
  echo "=== $name"
  timeout 300 python -m pytest tests/test_gpu_tc_layers.py "tests/test_gpu_parity.py::test_predict_parity_on_real_windows" "tests/test_gpu_parity.py::test_call_batch_goldens" -m gpu -x -q 2>&1 | tail -1
  timeout 120 python tools/tc_timeline.py 296 > gpurun_out/timeline_$name.txt 2>&1; tail -1 gpurun_out/timeline_$name.txt
  timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --no-configs 2>gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name', 'value %.0f large %.0f callk %.0f e2e %.0f parity %.2e' % (d['value'], d['config']['large_batch_reads_per_s'], d['config']['call_batch_kernel_reads_per_s_per_gpu'], d['e2e']['value'], d['parity']['max_abs_err']))
PY
done
