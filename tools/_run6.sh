set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
timeout 300 python tools/bench_postvar.py > gpurun_out/r2g_postvar.jsonl 2> gpurun_out/r2g_postvar.err
FGP_PV_NO_TMA=1 timeout 300 python tools/bench_postvar.py >> gpurun_out/r2g_postvar.jsonl 2>> gpurun_out/r2g_postvar.err
timeout 300 python tools/e2e_probe.py 20 > gpurun_out/r2g_e2e.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo bench rc=$?
tail -c 400 gpurun_out/r2g_bench.err
