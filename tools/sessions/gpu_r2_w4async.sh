#!/bin/bash
# third session of round 2: one-warp asynchronous instances -- parity tests, then the sweep over forms and instance counts
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "asynchronous or free_running" > $O/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2w_tests.log
timeout 600 python tools/async_forms.py > $O/r2w_forms.log 2>&1; echo "forms rc=$?"; cat $O/r2w_forms.log
