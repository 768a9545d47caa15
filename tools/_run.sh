O=gpurun_out
python -m pytest tests -x -q -m gpu -k "device_offsets" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/pcie_multi.py 2>&1 | grep -v "^\*\*\|OMP_NUM" | tee $O/pcie_n1.txt
