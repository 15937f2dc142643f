set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_bench_8gpu.json 2> gpurun_out/r2q_bench_8gpu.err; echo bench8 rc=$?
grep -c "NCCL INFO" gpurun_out/r2q_bench_8gpu.err; grep -m4 "nranks\|NVLS\|Init COMPLETE" gpurun_out/r2q_bench_8gpu.err | cut -c1-250
wc -l gpurun_out/r2q_bench_8gpu.json
tail -c 500 gpurun_out/r2q_bench_8gpu.json
