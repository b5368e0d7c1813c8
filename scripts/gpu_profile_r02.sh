# Round-2 evidence run (1 GPU): GPU tests, smoke, bench (default engine, fp32 engine, reference arm), ncu launch list +
# full captures of k_tc_forward (both forms) and k_merge_call, timelines, directory / fast5 benches.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/r02_nvidia_smi.csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.txt
cat gpurun_out/r02_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -4 gpurun_out/r02_smoke.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r02_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r02_bench_tcgen05.json
timeout 900 python bench.py --steps 10 --warmup 3 --engine fp32 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r02_bench_fp32.json
cut -c1-300 gpurun_out/r02_bench_tcgen05.json gpurun_out/r02_bench_fp32.json gpurun_out/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 300 --csv --log-file gpurun_out/r02_launches_tcgen05.csv python bench.py --steps 2 --warmup 3 --shard 8192 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_forward -s 8 -c 2 -o gpurun_out/r02_prof_tcgen05 python bench.py --steps 1 --warmup 3 --shard 2048 --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_merge_call -s 2 -c 1 -o gpurun_out/r02_prof_merge python tools/bench_callk.py > gpurun_out/r02_ncu_merge.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_forward -s 4 -c 1 -o gpurun_out/r02_prof_tcgen05_call python tools/bench_callk.py > gpurun_out/r02_ncu_call.log 2>&1
timeout 120 python tools/tc_timeline.py 296 > gpurun_out/r02_tc_timeline.txt 2>&1
timeout 120 python tools/tc_timeline.py 296 call > gpurun_out/r02_tc_timeline_call.txt 2>&1
timeout 300 python tools/bench_classify_dir.py 3000 2>&1 | tail -3 > gpurun_out/r02_classify_dir.txt
timeout 300 python tools/bench_fast5.py 2>&1 | tail -12 > gpurun_out/r02_fast5_reader.txt
ls -la gpurun_out | tail -30
