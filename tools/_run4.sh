set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
R=$GRAFT_REPO_ROOT/fastgaussianprocesses_b200
timeout 300 python tools/coop_probe.py > gpurun_out/r2e_probe.jsonl 2> gpurun_out/r2e_probe.err
FGP_CAP_C=11 FGP_COLS_LOG2=2 timeout 300 python tools/coop_probe.py >> gpurun_out/r2e_probe.jsonl 2>> gpurun_out/r2e_probe.err
FGP_CAP_C=11 FGP_COLS_LOG2=3 timeout 300 python tools/coop_probe.py >> gpurun_out/r2e_probe.jsonl 2>> gpurun_out/r2e_probe.err
FGP_B200_LIB=$R/lib_rv_t512/libfgp_b200.so timeout 300 python tools/coop_probe.py >> gpurun_out/r2e_probe.jsonl 2>> gpurun_out/r2e_probe.err
FGP_COLS_LOG2=2 FGP_B200_LIB=$R/lib_rv_t512/libfgp_b200.so timeout 300 python tools/coop_probe.py >> gpurun_out/r2e_probe.jsonl 2>> gpurun_out/r2e_probe.err
timeout 300 python tools/bench_postvar.py > gpurun_out/r2e_postvar.jsonl 2> gpurun_out/r2e_postvar.err
timeout 300 python tools/e2e_probe.py 20 > gpurun_out/r2e_e2e.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo bench rc=$?
tail -c 300 gpurun_out/r2e_bench.err
