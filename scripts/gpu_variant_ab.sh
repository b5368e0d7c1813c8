# ENGINE=tcgen05-split bash scripts/gpu_variant_ab.sh ... benches another engine than the default.
# A/B of kernel variants: for every variants/*.so run the tcgen05 parity tests (quick subset), the
# timeline and a short bench.  Usage: bash scripts/gpu_variant_ab.sh [variant.so ...]
mkdir -p gpurun_out
libs="$@"
[ -z "$libs" ] && libs="deepbinner_b200/libdeepbinner_b200.so $(ls variants/*.so 2>/dev/null)"
for lib in $libs; do
  name=$(basename $lib .so)
  export DEEPBINNER_B200_LIB=$PWD/$lib
  echo "=== $name"
  timeout 600 python -m pytest tests/test_gpu_tc_layers.py "tests/test_gpu_parity.py::test_predict_parity_on_real_windows" "tests/test_gpu_parity.py::test_split_engine_parity" -m gpu -x -q 2>&1 | tail -1
  timeout 300 python tools/tc_timeline.py 296 > gpurun_out/timeline_$name.txt 2>&1; tail -1 gpurun_out/timeline_$name.txt
  timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 2 ${ENGINE:+--engine $ENGINE} 2>gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name', 'value %.0f e2e %.0f frac %.4f parity %.2e clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['max_abs_err'], d['clocks']))
PY
done
