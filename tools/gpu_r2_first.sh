#!/bin/bash
# Round-2 first GPU session: topology facts, smoke, GPU parity tests, PCIe probe.
TAG=${1:-r2a}
OUT=gpurun_out; mkdir -p $OUT
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current,pci.bus_id --format=csv
  nvidia-smi topo -m
  lscpu | head -30
  nproc
  numactl -H 2>&1 | head -20
  for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist 2>/dev/null); fi; done
  free -g
} > $OUT/topo_$TAG.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke_$TAG.log
timeout 1700 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -30 $OUT/pytest_gpu_$TAG.log
python - > $OUT/probe_$TAG.txt 2>&1 <<'PY'
from p265_b200.engine import Engine
e = Engine(0)
for sz in (32<<20, 256<<20):
    print("both", sz>>20, [round(v/1e9,2) for v in e.pcie_probe(sz, 6)])
    print("h2d only", sz>>20, round(e.pcie_probe(sz, 6, d2h=False)[0]/1e9,2))
    print("d2h only", sz>>20, round(e.pcie_probe(sz, 6, h2d=False)[1]/1e9,2))
PY
cat $OUT/probe_$TAG.txt
