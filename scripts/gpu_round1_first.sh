set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -30
python bench.py --steps 5 --warmup 3 --cpu-seconds 5 2>&1 | tail -3 > gpurun_out/bench_fp32_first.json
cat gpurun_out/bench_fp32_first.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --steps 1 --warmup 3 --shard 2048 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/launches_fp32.csv
ncu --set full --clock-control none --import-source on -k regex:k_fp32_predict -s 8 -c 2 -o gpurun_out/prof_fp32 python bench.py --steps 1 --warmup 3 --shard 2048 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
