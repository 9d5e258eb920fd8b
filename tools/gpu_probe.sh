#!/bin/bash
# Host/link topology of the GPU box: what the end-to-end path can expect (NUMA nodes, PCIe placement, cpuset).
OUT=${1:-gpurun_out}
mkdir -p $OUT
{
  echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm --format=csv
  echo "== topo"; nvidia-smi topo -m 2>&1 | head -30
  echo "== lscpu"; lscpu | egrep "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
  echo "== nodes"; ls /sys/devices/system/node/ 2>&1 | tr '\n' ' '; echo
  for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== allowed"; grep -i allowed /proc/self/status
  echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [[ "$(cat $d/class)" == 0x0302* ]]; then echo "$d numa=$(cat $d/numa_node) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
  echo "== nproc"; nproc; free -g | head -2
  echo "== thp"; cat /sys/kernel/mm/transparent_hugepage/enabled
} > $OUT/box_probe.txt 2>&1
cat $OUT/box_probe.txt
