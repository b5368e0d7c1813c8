set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 2>&1 | tail -1 > gpurun_out/bench_tc_v1.json
cat gpurun_out/bench_tc_v1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_forward -s 6 -c 1 -o gpurun_out/prof_tc_v1 python bench.py --steps 1 --warmup 3 --shard 2048 --no-cpu-baseline > gpurun_out/ncu_tc_v1.log 2>&1
ls -la gpurun_out
