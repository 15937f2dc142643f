python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | cut -c1-300 | tail -5
for cfg in "" "FGP_THREAD_MUL=2" "FGP_CAP_C=11 FGP_THREAD_MUL=2 FGP_COLS_LOG2=2" "FGP_NO_HS=1"; do
  env $cfg python tools/tune_mll.py 20 8 lattice
done 2>&1 | tee gpurun_out/tune_hs3.jsonl
for cfg in "" "FGP_THREAD_MUL=2" "FGP_CAP_C=11 FGP_THREAD_MUL=2 FGP_COLS_LOG2=2"; do
  env $cfg python tools/tune_mll.py 18 8 lattice
  env $cfg python tools/tune_mll.py 16 4 net
  env $cfg python tools/tune_mll.py 20 8 net
done 2>&1 | tee -a gpurun_out/tune_hs3.jsonl
