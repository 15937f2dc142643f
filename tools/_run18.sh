set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/debug_mb.py 2>&1 | tail -40
