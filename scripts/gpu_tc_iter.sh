# one iteration of tcgen05 engine work: per-layer check, parity tests, timeline, short bench
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_tc_layers.py -m gpu -x -q -s 2>&1 | tail -50 > gpurun_out/tc_layers.txt; tail -4 gpurun_out/tc_layers.txt
grep -q "1 passed" gpurun_out/tc_layers.txt || { grep -i "error\|fail\|trap" gpurun_out/tc_layers.txt | head; exit 1; }
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/tc_timeline.py 296 > gpurun_out/tc_timeline.txt 2>&1; tail -1 gpurun_out/tc_timeline.txt
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 2 2>gpurun_out/bench_iter.err | tail -1 > gpurun_out/bench_tc_iter.json
cut -c1-200 gpurun_out/bench_tc_iter.json
