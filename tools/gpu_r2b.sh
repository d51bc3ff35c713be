#!/bin/bash
# round 2, call B (2 GPUs): peer-memory paths for real (cudaIpc + NVLink), then the bench at N=2
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/r2b_topo.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check_peer.py 1000000 1000000 100000 10 > $out/r2b_peer2.log 2>&1; echo "dist_check_peer rc=$?"
grep -v "^W\|^\*\*\*" $out/r2b_peer2.log | tail -25
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $out/r2b_bench2.json 2> $out/r2b_bench2.err; echo "bench2 rc=$?"
cat $out/r2b_bench2.json; tail -5 $out/r2b_bench2.err
