#!/bin/bash
O=gpurun_out
{ nvidia-smi topo -m; numactl -H; lscpu | head -25; echo cpus; cat /sys/fs/cgroup/cpuset.cpus.effective; echo mems; cat /sys/fs/cgroup/cpuset.mems.effective; nproc; ls /sys/devices/system/node/; nvidia-smi --query-gpu=pci.bus_id --format=csv; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node); fi; done; } > $O/topo_r02.txt 2>&1
for c in 6 8 12 16 24; do GGP_B200_UPLOAD_CHUNKS=$c python tools/r02_jobs/e2e_probe.py fast 20; done > $O/e2e_chunks_r02.txt 2>&1
GGP_B200_TIMELINE=1 python tools/r02_jobs/e2e_probe.py fast 3 > $O/e2e_timeline_fast_r02.txt 2>&1
python tools/fast_probe.py 10000 0,4,5,6 10 > $O/fast_probe_r02.txt 2>&1
GGP_B200_FAST_CHUNKED=0 python tools/fast_probe.py 10000 5 10 >> $O/fast_probe_r02.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_fast5_r02.csv python tools/fast_probe.py 10000 5 1 > $O/ncu_l_fast5_r02.log 2>&1
GGP_B200_FAST_CHUNKED=0 ncu --set full --clock-control none --import-source on -k regex:ggp_fast_loglik --launch-skip 17 --launch-count 1 -f -o $O/prof_fast5_r02 python tools/fast_probe.py 10000 5 1 > $O/ncu_f_fast5_r02.log 2>&1
cat $O/e2e_chunks_r02.txt $O/fast_probe_r02.txt
