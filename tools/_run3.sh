set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
timeout 300 python tools/bench_fwht_fused.py > gpurun_out/r2d_fwht_fused.json 2> gpurun_out/r2d_fwht_fused.err; echo rc=$?
timeout 300 python tools/coop_probe.py > gpurun_out/r2d_probe.json 2> gpurun_out/r2d_probe.err; echo rc=$?
for c in 1 2 4 8 16; do FGP_PV_CHUNK=$c timeout 300 python tools/bench_postvar.py >> gpurun_out/r2d_postvar.jsonl 2>> gpurun_out/r2d_postvar.err; done
FGP_B200_NO_PVZ=1 timeout 300 python tools/bench_postvar.py >> gpurun_out/r2d_postvar.jsonl 2>> gpurun_out/r2d_postvar.err
timeout 300 python tools/e2e_probe.py 20 > gpurun_out/r2d_e2e_three.txt 2>&1
FGP_COOP=1 FGP_CAP_C=11 FGP_COLS_LOG2=2 timeout 300 python tools/e2e_probe.py 20 > gpurun_out/r2d_e2e_coop.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo bench rc=$?
python tools/profile_target.py pvarz 1 > gpurun_out/r2d_plain_pvarz.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pv_pass -c 4 -o gpurun_out/r2d_ncu_pvarz python tools/profile_target.py pvarz 1 > gpurun_out/r2d_ncu_pvarz.log 2>&1
python tools/profile_target.py mllz 2 > gpurun_out/r2d_plain_mllz.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mll_pass -s 3 -c 3 -o gpurun_out/r2d_ncu_mllz python tools/profile_target.py mllz 2 > gpurun_out/r2d_ncu_mllz.log 2>&1
ls -la gpurun_out | tail -20
