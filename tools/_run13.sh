set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -3 gpurun_out/r2n_pytest.log
timeout 300 python tools/bench_postvar.py > gpurun_out/r2n_postvar.json 2> gpurun_out/r2n_postvar.err
FGP_B200_NO_PVZ=1 timeout 300 python tools/bench_postvar.py > gpurun_out/r2n_postvar_unfused.json 2>> gpurun_out/r2n_postvar.err
timeout 300 python tools/e2e_probe.py 20 > gpurun_out/r2n_e2e.txt 2>&1
