set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "data_spectrum" 2>&1 | tail -5
timeout 1500 python tools/parity_report.py > gpurun_out/PARITY_r02.json 2> gpurun_out/parity.err; echo rc=$?; tail -3 gpurun_out/parity.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/PARITY_r02.json"))
print(json.dumps(d["worst"],indent=0)[:3000])
PY
