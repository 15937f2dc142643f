set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err; echo bench2 rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_ref_2gpu.json 2> gpurun_out/r2l_bench_ref_2gpu.err; echo ref2 rc=$?
grep -c "NCCL INFO" gpurun_out/r2l_bench_2gpu.err; grep -m3 "nranks\|Init COMPLETE" gpurun_out/r2l_bench_2gpu.err | cut -c1-250
tail -c 600 gpurun_out/r2l_bench_2gpu.json
