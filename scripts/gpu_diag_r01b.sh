set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 120 ./tools/tc_microbench > gpurun_out/tc_microbench.txt 2>&1; tail -5 gpurun_out/tc_microbench.txt
timeout 300 python tools/tc_timeline.py 296 > gpurun_out/tc_timeline_fine.txt 2>&1; tail -3 gpurun_out/tc_timeline_fine.txt
