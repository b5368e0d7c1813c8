set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_layers.py -m gpu -x -q -s 2>&1 | tail -60
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
