cd $GRAFT_REPO_ROOT
for v in 0 1; do FGP_PDL=$v timeout 300 python tools/tune_mll.py 20 8 lattice 2>&1 | tail -1 | cut -c1-200; done
for v in 0 1; do FGP_PDL=$v timeout 300 python tools/tune_mll.py 18 8 lattice 2>&1 | tail -1 | cut -c1-200; done
for v in 0 1; do FGP_PDL=$v timeout 300 python tools/tune_mll.py 16 4 net 2>&1 | tail -1 | cut -c1-200; done
