#!/bin/bash
# Dump what the host side of the e2e path depends on: CPU quota/affinity, NUMA nodes, GPU <-> NUMA mapping.
echo "== nproc / affinity"; nproc; taskset -p $$ 2>/dev/null; grep Cpus_allowed_list /proc/self/status
echo "== cgroup"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
echo "== lscpu"; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|^CPU\(s\)|L3|Flags" | sed 's/Flags.*avx512/Flags: ... avx512/' | cut -c1-300
echo "== numa nodes"; for n in /sys/devices/system/node/node*; do echo "$n cpulist=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "== gpus"; nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv,noheader
for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do
  bb=$(echo $b | tr 'A-Z' 'a-z' | sed 's/^0000//'); p=/sys/bus/pci/devices/$bb
  echo "$b numa_node=$(cat $p/numa_node 2>/dev/null) local_cpulist=$(cat $p/local_cpulist 2>/dev/null) link=$(cat $p/current_link_speed 2>/dev/null) x$(cat $p/current_link_width 2>/dev/null)"
done
echo "== topo"; nvidia-smi topo -m 2>/dev/null
echo "== mem policy syscalls"; python - <<'PY'
import ctypes, os
libc = ctypes.CDLL(None, use_errno=True)
mode = ctypes.c_int(-1)
r = libc.syscall(239, ctypes.byref(mode), None, 0, None, 0)   # get_mempolicy
print("get_mempolicy rc", r, "mode", mode.value, "errno", ctypes.get_errno())
PY
