#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python tools/group_time.py 2 2000000 > $out/r2g_group_time.log 2>&1; echo "group_time rc=$?"; grep -v "^\[vpc group\]" $out/r2g_group_time.log | tail -5; grep "^\[vpc group\]" $out/r2g_group_time.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check_peer.py 1000000 1000000 100000 10 > $out/r2g_peer2.log 2>&1; echo "dist_check_peer rc=$?"
grep "rank 0" $out/r2g_peer2.log | tail -12
