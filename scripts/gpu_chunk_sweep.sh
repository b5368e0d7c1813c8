# sweep of the call_batch pipeline's chunk size (network windows per launch of the persistent kernel)
mkdir -p gpurun_out
for c in 1024 2048 4096 8192 16384; do
  DEEPBINNER_B200_CALL_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 2>gpurun_out/sweep_$c.err | tail -1 > gpurun_out/sweep_$c.json
  python - <<PY
import json
d=json.load(open('gpurun_out/sweep_$c.json'))
c=d['config']['configs']
print('chunk %6d  value %.3f M  e2e %.3f M  cfg2 %.3f Mw/s  cfg3 %.3f Mw/s  cfg5 %.3f Mw/s' % ($c, d['value']/1e6, d['e2e']['value']/1e6, c['native_start_end_batch256']['windows_per_s']/1e6, c['rapid_start_batch512']['windows_per_s']/1e6, c['realtime_stream_start_end']['windows_per_s']/1e6))
PY
done
