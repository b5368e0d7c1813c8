set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/tc_timeline.py 296 2>&1 | tail -50
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 2>&1 | tail -1 > gpurun_out/bench_tc_iter.json
cut -c1-420 gpurun_out/bench_tc_iter.json
