# auto chunk size (default) against the fixed 2048-window chunks, interleaved twice
mkdir -p gpurun_out
for rep in 1 2; do
for c in auto 2048; do
  if [ $c = auto ]; then unset DEEPBINNER_B200_CALL_CHUNK; else export DEEPBINNER_B200_CALL_CHUNK=$c; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 2>gpurun_out/sweep_$c.err | tail -1 > gpurun_out/sweep_$c.json
  python - <<PY
import json
d=json.load(open('gpurun_out/sweep_$c.json'))
c=d['config']['configs']
print('chunk %6s  value %.3f M  e2e %.3f M  cfg2 %.3f Mw/s  cfg3 %.3f Mw/s  cfg5 %.3f Mw/s' % ('$c', d['value']/1e6, d['e2e']['value']/1e6, c['native_start_end_batch256']['windows_per_s']/1e6, c['rapid_start_batch512']['windows_per_s']/1e6, c['realtime_stream_start_end']['windows_per_s']/1e6))
PY
done
done
unset DEEPBINNER_B200_CALL_CHUNK
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
