#!/bin/bash
# ncu of a multimem.st (NVLS multicast) launch: the single-process probe's k_mc_simt over a 2-GPU multicast group.
# Kernel replay cannot save / restore multicast-mapped memory ("Failed to profile"), so the whole (short) program is replayed.
OUT=gpurun_out/r2/ncu_mc
mkdir -p $OUT
timeout 150 bin/bb-p2p-probe --gpus 2 --only-mc --mib 512 --iters 2 > $OUT/probe_only_mc.txt 2>&1
timeout 800 ncu --replay-mode application --set full --section Nvlink_Tables --section Nvlink_Topology --import-source on --clock-control none \
  -k regex:k_mc_simt -c 1 -f -o $OUT/mc_simt_n2 bin/bb-p2p-probe --gpus 2 --only-mc --mib 512 --iters 2 > $OUT/probe_under_ncu.txt 2> $OUT/ncu.err
ncu -i $OUT/mc_simt_n2.ncu-rep --page raw --csv > $OUT/mc_simt_n2_raw.csv 2>/dev/null
ls -la $OUT; tail -5 $OUT/ncu.err; tail -3 $OUT/probe_under_ncu.txt | cut -c1-300; cat $OUT/probe_only_mc.txt | cut -c1-250
