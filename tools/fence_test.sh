cd $GRAFT_REPO_ROOT
echo "== no fence"; SE2GPU_BA_DEBUG=1 python tools/ba_debug.py 2>&1 | grep "phase schur \|phase linearize\|phase backsub" | tail -3
echo "== sys fence"; SE2GPU_BA_DEBUG_SYSFENCE=1 SE2GPU_BA_DEBUG=1 python tools/ba_debug.py 2>&1 | grep "phase schur \|phase linearize\|phase backsub" | tail -3
