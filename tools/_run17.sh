set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multitask_gpu.py -q -m gpu -k "batched or guards or nugget" 2>&1 | grep -v "^    \|^$" | tail -60
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
