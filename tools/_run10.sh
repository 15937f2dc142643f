set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
timeout 300 python tools/bench_net.py > gpurun_out/r2k_net.json 2> gpurun_out/r2k_net.err
timeout 300 python tools/e2e_probe.py 20 > gpurun_out/r2k_e2e.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo bench2 rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_ref_2gpu.json 2> gpurun_out/r2k_bench_ref_2gpu.err; echo ref2 rc=$?
tail -c 1500 gpurun_out/r2k_bench_2gpu.err
