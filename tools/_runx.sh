cd $GRAFT_REPO_ROOT
for cap in 0 9 10 11; do FGP_CAP_R=$cap timeout 300 python tools/tune_mll.py 16 4 net 2>&1 | tail -1 | cut -c1-130; done
for cap in 0 10 11; do FGP_CAP_R=$cap timeout 300 python tools/tune_mll.py 18 8 net 2>&1 | tail -1 | cut -c1-130; done
for cap in 0 11; do FGP_CAP_R=$cap timeout 300 python tools/tune_mll.py 20 8 net 2>&1 | tail -1 | cut -c1-130; done
for cap in 0 9 10; do FGP_CAP_C=$cap timeout 300 python tools/tune_mll.py 16 8 lattice 2>&1 | tail -1 | cut -c1-130; done
for cap in 0 10; do FGP_CAP_C=$cap timeout 300 python tools/tune_mll.py 18 8 lattice 2>&1 | tail -1 | cut -c1-130; done
