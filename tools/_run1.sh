set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python tools/coop_probe.py > gpurun_out/r2a_probe_main.json 2> gpurun_out/r2a_probe_main.err; echo rc=$?
FGP_B200_LIB=$GRAFT_REPO_ROOT/fastgaussianprocesses_b200/lib_rt/libfgp_b200.so timeout 300 python tools/coop_probe.py --skip-three > gpurun_out/r2a_probe_stamps.json 2> gpurun_out/r2a_probe_stamps.err; echo rc=$?
for cfg in "FGP_COOP_CTAS=148" "FGP_CAP_C=11 FGP_COLS_LOG2=2" "FGP_CAP_C=11 FGP_COLS_LOG2=3" "FGP_CAP_C=11 FGP_COLS_LOG2=2 FGP_COOP_CTAS=148" "FGP_COLS_LOG2=2" "FGP_COLS_LOG2=4" "FGP_CAP_C=10 FGP_COLS_LOG2=2"; do
  env $cfg FGP_B200_LIB=$GRAFT_REPO_ROOT/fastgaussianprocesses_b200/lib_rt/libfgp_b200.so timeout 200 python tools/coop_probe.py --skip-three >> gpurun_out/r2a_sweep.jsonl 2>> gpurun_out/r2a_sweep.err
done
timeout 200 python tools/coop_probe.py --log2n 18 --B 64 --iters 40 --chunk 10 >> gpurun_out/r2a_probe_other.jsonl 2>> gpurun_out/r2a_probe_other.err
timeout 200 python tools/coop_probe.py --log2n 16 --d 4 --family dnb2 >> gpurun_out/r2a_probe_other.jsonl 2>> gpurun_out/r2a_probe_other.err
timeout 200 python tools/coop_probe.py --log2n 22 --d 8 --iters 50 --chunk 10 >> gpurun_out/r2a_probe_other.jsonl 2>> gpurun_out/r2a_probe_other.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo bench rc=$?
cat gpurun_out/r2a_probe_main.json
