set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
timeout 300 python tools/bench_postvar.py > gpurun_out/r2o_postvar.json 2> gpurun_out/r2o_postvar.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo bench rc=$?
python tools/profile_target.py pvarz 1 > gpurun_out/r2o_plain_pvarz.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pv_pass -c 4 -o gpurun_out/r2o_ncu_pvarz python tools/profile_target.py pvarz 1 > gpurun_out/r2o_ncu_pvarz.log 2>&1
