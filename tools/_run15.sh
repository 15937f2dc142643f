set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -14 gpurun_out/r2p_pytest.log
python tools/profile_target.py netpm 1 > gpurun_out/r2p_plain_netpm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:post_mean_kernel -c 1 -o gpurun_out/r2p_ncu_netpm python tools/profile_target.py netpm 1 > gpurun_out/r2p_ncu_netpm.log 2>&1
python tools/profile_target.py netpv 1 > gpurun_out/r2p_plain_netpv.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pvn_pass -c 2 -o gpurun_out/r2p_ncu_netpv python tools/profile_target.py netpv 1 > gpurun_out/r2p_ncu_netpv.log 2>&1
