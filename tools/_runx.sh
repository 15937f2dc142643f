cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/tune_mll.py 20 8 lattice 2>&1 | tail -1 | cut -c1-420
timeout 300 python tools/tune_mll.py 20 8 net 2>&1 | tail -1 | cut -c1-300
FGP_B200_LIB=$PWD/fastgaussianprocesses_b200/lib_rt/libfgp_b200.so timeout 300 python tools/pass_stamps.py 20 8 > gpurun_out/r2w_pass_stamps_n20.json 2> gpurun_out/r2w_err.log; tail -3 gpurun_out/r2w_err.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2w_pass_stamps_n20.json"))
print(d["ctas"], d["kernel_span_us"], d["iteration_span_us_5runs"], d.get("fit_tail_us"))
for k,v in d["phase_us"].items(): print("      %-18s"%k, v)
PY
