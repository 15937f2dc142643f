set -x
cd $GRAFT_REPO_ROOT
for v in "" _lb3 _lb4; do
  FGP_B200_LIB=$PWD/fastgaussianprocesses_b200/lib$v/libfgp_b200.so timeout 300 python tools/bench_postvar.py 2>&1 | tail -1
  FGP_B200_LIB=$PWD/fastgaussianprocesses_b200/lib$v/libfgp_b200.so timeout 300 python tools/tune_mll.py 20 8 lattice 2>&1 | tail -1 | cut -c1-400
done
