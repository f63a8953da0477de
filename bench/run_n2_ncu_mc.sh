#!/bin/bash
# ncu of a multimem.st (NVLS multicast) launch: the single-process probe's k_mc_simt over a 2-GPU multicast group.
OUT=gpurun_out/r2/ncu_mc
mkdir -p $OUT
timeout 600 ncu --set full --section Nvlink_Tables --section Nvlink_Topology --import-source on --clock-control none -k regex:k_mc_simt -c 2 -f -o $OUT/mc_simt_n2 \
  bin/bb-p2p-probe --gpus 2 --quick --mib 512 --iters 2 > $OUT/probe_under_ncu.txt 2> $OUT/ncu.err
ncu -i $OUT/mc_simt_n2.ncu-rep --page raw --csv > $OUT/mc_simt_n2_raw.csv 2>/dev/null
ls -la $OUT; tail -5 $OUT/ncu.err; grep -c "" $OUT/mc_simt_n2_raw.csv
