#!/bin/bash
# 2-GPU session: the library's own NCCL communicator -- python (torchrun) parity vs the oracle, and the CLI driver with TIKTAK_NGPU=2
cd "$GRAFT_REPO_ROOT" || exit 1
O=$PWD/gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > $O/r2g_mgpu_w2.log 2>&1; echo "mgpu_check rc=$?" > $O/r2g_status.txt
# driver: Griewank 10-D, 100k points, 95 restarts, rounds of 16; 1 GPU vs 2 GPUs -> identical .dat files
mk() { mkdir -p $1; cd $1; { echo 10; echo 1; echo -1; echo 100000; echo -1; echo -1; echo 95; echo 0; echo -1; for i in $(seq 10); do echo "-400.0, 600.0"; done; for i in $(seq 10); do echo 100.0; done; for i in $(seq 10); do echo "-400.0, 600.0"; done; } > config.txt; }
(mk /tmp/run1; TIKTAK_OBJECTIVE=1 TIKTAK_ROUND_SIZE=16 timeout 300 $GRAFT_REPO_ROOT/driver/TiktakGlobalSearch 0 config.txt a > out.txt 2>&1; echo "driver ngpu1 rc=$?" >> $O/r2g_status.txt)
(mk /tmp/run2; TIKTAK_NGPU=2 TIKTAK_OBJECTIVE=1 TIKTAK_ROUND_SIZE=16 timeout 300 $GRAFT_REPO_ROOT/driver/TiktakGlobalSearch 0 config.txt a > out.txt 2>&1; echo "driver ngpu2 rc=$?" >> $O/r2g_status.txt)
for f in sobolFnVal.dat x_starts.dat searchStart.dat searchResults.dat FinalResults.dat legitSobol.dat; do cmp /tmp/run1/$f /tmp/run2/$f >> $O/r2g_status.txt 2>&1 && echo "same $f ($(wc -l < /tmp/run1/$f) lines)" >> $O/r2g_status.txt; done
tail -3 /tmp/run2/out.txt >> $O/r2g_status.txt; tail -2 /tmp/run2/logFile.txt >> $O/r2g_status.txt
tail -6 $O/r2g_mgpu_w2.log; cat $O/r2g_status.txt
