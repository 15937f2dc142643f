set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
timeout 300 python tools/coop_probe.py > gpurun_out/r2f_probe.jsonl 2> gpurun_out/r2f_probe.err
timeout 300 python tools/coop_probe.py --log2n 18 --B 64 --iters 40 --chunk 10 >> gpurun_out/r2f_probe.jsonl 2>> gpurun_out/r2f_probe.err
timeout 300 python tools/coop_probe.py --log2n 22 --iters 50 --chunk 10 >> gpurun_out/r2f_probe.jsonl 2>> gpurun_out/r2f_probe.err
timeout 300 python tools/e2e_probe.py 20 > gpurun_out/r2f_e2e.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo bench rc=$?
timeout 1200 python tools/parity_report.py > gpurun_out/PARITY.json 2> gpurun_out/r2f_parity.err; echo parity rc=$?
