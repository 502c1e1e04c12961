#!/bin/bash
# asynchronous instances: parity tests against the oracle's event simulation, then C2 / C3 restart stage, rounds against instances
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "asynchronous or local_min" > $O/r2async_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/r2async_tests.log
timeout 600 python tools/async_time.py 2>&1 | tail -20
