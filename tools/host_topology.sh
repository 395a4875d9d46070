#!/bin/bash
# Host-side topology of a GPU box: what the end-to-end (host-pointer) path has to live with.
echo "== nproc"; nproc
echo "== lscpu"; lscpu | egrep -i "model name|socket|numa|^cpu\(s\)|thread|core\(s\)|hypervisor|virtual" 
echo "== nodes"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "== topo"; nvidia-smi topo -m 2>&1
echo "== gpu pci"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [[ "$(cat $d/class)" == 0x0302* ]]; then echo "$d numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist)"; fi; done
echo "== affinity of this shell"; taskset -p $$ 2>&1; grep -i cpus_allowed_list /proc/self/status
echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
echo "== mem"; free -g | head -2
which numactl 2>&1
