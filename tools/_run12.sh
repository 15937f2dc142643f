set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -3 gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo ref rc=$?
timeout 300 python tools/bench_transforms.py > gpurun_out/r2m_transforms.json 2> gpurun_out/r2m_transforms.err
timeout 300 python tools/bench_postvar.py > gpurun_out/r2m_postvar.json 2> gpurun_out/r2m_postvar.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --log2m 13 --log2mv 9 > gpurun_out/r2m_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2m_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --log2m 13 --log2mv 9 > gpurun_out/r2m_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo smoke rc=$?
