#!/bin/bash
# Scaling evidence on N GPUs of one box: the bench line (weak + strong curves, e2e, bare-copy host ceiling) and the
# PCIe probe.  Usage (on the GPU box): bash tools/scale_runs.sh N
N=$1
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > $O/final_bench_n$N.json 2> $O/final_bench_n$N.err
tail -c 300 $O/final_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_multi.py 2>&1 | grep "ranks=" > $O/final_pcie_n$N.txt
cat $O/final_pcie_n$N.txt
