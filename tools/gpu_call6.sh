#!/bin/bash
mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done; python -c "
import torch
for i in range(torch.cuda.device_count()):
    p=torch.cuda.get_device_properties(i); print(i, p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
"; numactl -H 2>/dev/null | head -20 ) > gpurun_out/r02_topology.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_handoff.py -m gpu -x -q -s > gpurun_out/r02_handoff_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_handoff_tests.log
grep -E "handoff|forward from|passed|failed|Error|error|rc=" gpurun_out/r02_handoff_tests.log | head -30
cat gpurun_out/r02_topology.txt | head -60
