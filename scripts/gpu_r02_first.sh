# Round 2, first GPU call: per-layer check + parity of the solo kernel, timelines, A/B bench pair vs solo.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_layers.py -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r02_layers.txt
tail -5 gpurun_out/r02_layers.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "unsaturated or real_windows or goldens or edge or ragged" 2>&1 | tail -40 > gpurun_out/r02_parity_first.txt
tail -8 gpurun_out/r02_parity_first.txt
timeout 300 python tools/tc_timeline.py 296 tcgen05 > gpurun_out/r02_timeline_solo.txt 2>&1; tail -3 gpurun_out/r02_timeline_solo.txt
timeout 300 python tools/tc_timeline.py 1 tcgen05 > gpurun_out/r02_timeline_solo_1cta.txt 2>&1; tail -1 gpurun_out/r02_timeline_solo_1cta.txt
for eng in tcgen05-pair tcgen05; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --engine $eng 2>gpurun_out/r02_bench_$eng.err | tail -1 > gpurun_out/r02_bench_$eng.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_$eng.json'))
print('$eng', 'value %.0f e2e %.0f big %.0f frac %.4f parity %.2e clocks %s' % (d['value'], d['e2e']['value'], d['config']['large_batch_reads_per_s'], d['roofline']['frac'], d['parity']['max_abs_err'], d['clocks']))
PY
done
