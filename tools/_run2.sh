set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
R=$GRAFT_REPO_ROOT/fastgaussianprocesses_b200
for v in lib_rt lib_rv_twf lib_rv_minb1 lib_rv_noinl lib_rv_twf_noinl; do
  FGP_B200_LIB=$R/$v/libfgp_b200.so timeout 200 python tools/coop_probe.py --skip-three >> gpurun_out/r2c_variants.jsonl 2>> gpurun_out/r2c_variants.err
  FGP_CAP_C=11 FGP_COLS_LOG2=2 FGP_B200_LIB=$R/$v/libfgp_b200.so timeout 200 python tools/coop_probe.py --skip-three >> gpurun_out/r2c_variants.jsonl 2>> gpurun_out/r2c_variants.err
done
timeout 300 python tools/bench_fwht_fused.py > gpurun_out/r2c_fwht_fused.json 2> gpurun_out/r2c_fwht_fused.err; echo rc=$?
FGP_FUSED_NO_TMA=1 timeout 300 python tools/bench_fwht_fused.py > gpurun_out/r2c_fwht_fused_notma.json 2>> gpurun_out/r2c_fwht_fused.err; echo rc=$?
FGP_COOP=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; echo ref rc=$?
FGP_COOP=0 timeout 1200 python tools/parity_report.py > gpurun_out/PARITY.json 2> gpurun_out/r2c_parity.err; echo parity rc=$?
tail -c 600 gpurun_out/r2c_bench.json
