#!/bin/bash
# What does this GPU box give?  cores, memory, disks, GPU topology, file-write rates (tools/io_probe.cpp),
# host transpose rate (tools/host_mirror_exp.cpp).  Output goes to gpurun_out/box_probe.log.
out=gpurun_out/box_probe.log
mkdir -p gpurun_out
{
echo "== cpu"; nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|L3|MHz" 
echo "== memory"; free -g
echo "== filesystems"; df -h . /dev/shm /tmp 2>/dev/null
echo "== gpus"; nvidia-smi --query-gpu=index,name,memory.total,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
nvidia-smi topo -m 2>/dev/null | head -30
echo "== io_probe /dev/shm"; tools/_bin/io_probe /dev/shm ${1:-12} $(nproc)
echo "== io_probe disk ($PWD/gpurun_out)"; tools/_bin/io_probe $PWD/gpurun_out ${1:-12} $(nproc)
echo "== host mirror"; for m in 0 1 2; do tools/_bin/host_mirror_exp 32768 $(nproc) $m; done
} > $out 2>&1
tail -50 $out
