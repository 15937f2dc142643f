set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/e2e_timeline.py 20 2>&1 | grep -v Warn | grep "host us\|GPU span"
timeout 300 python tools/fit_trace.py 20 2>&1 | grep -v Warn | grep "sync points" | cut -c1-700
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2s_bench_line.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s_bench_line.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","warm") if k in d})
PY
