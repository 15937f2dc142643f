set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/fit_trace.py 20 2>&1 | grep -v Warn | grep "timeline\|sync points" | cut -c1-1500
timeout 300 python tools/e2e_hostprof.py 20 2>&1 | grep -v Warn | tail -52 | cut -c1-160
