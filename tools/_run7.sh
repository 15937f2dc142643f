set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
timeout 300 python tools/bench_postvar.py > gpurun_out/r2h_postvar.jsonl 2> gpurun_out/r2h_postvar.err
FGP_PV_TMA=0 timeout 300 python tools/bench_postvar.py >> gpurun_out/r2h_postvar.jsonl 2>> gpurun_out/r2h_postvar.err
FGP_COLS_LOG2=4 timeout 300 python tools/bench_postvar.py >> gpurun_out/r2h_postvar.jsonl 2>> gpurun_out/r2h_postvar.err
FGP_COLS_LOG2=4 FGP_PV_TMA=0 timeout 300 python tools/bench_postvar.py >> gpurun_out/r2h_postvar.jsonl 2>> gpurun_out/r2h_postvar.err
timeout 300 python tools/fit_trace.py 20 > gpurun_out/r2h_fit_trace.txt 2>&1
