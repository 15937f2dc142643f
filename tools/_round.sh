python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -x -q 2>&1 | grep -v Warning | cut -c1-300 | tail -3
python tools/tune_mll.py 20 8 lattice; python tools/tune_mll.py 18 8 lattice; python tools/tune_mll.py 20 8 net; python tools/tune_mll.py 16 4 net
