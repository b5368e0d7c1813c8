# final round-2 consistency run: GPU suite three times (flakiness), smoke, fast5 reader bench, default bench line
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest_$i.txt; tail -1 gpurun_out/final_pytest_$i.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; tail -2 gpurun_out/final_smoke.txt
timeout 600 python tools/bench_fast5.py > gpurun_out/final_fast5_reader.txt 2>&1; head -20 gpurun_out/final_fast5_reader.txt
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json
cut -c1-400 gpurun_out/final_bench.json
