mkdir -p gpurun_out/r2/ncu
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench_n2c.json 2> gpurun_out/r2/bench_n2c.err
timeout 200 ncu --set full --section Nvlink_Tables --section Nvlink_Topology --section Nvlink --clock-control none --import-source on -k regex:bb_xfer -s 0 -c 4 -o gpurun_out/r2/ncu/xfer_xxh3_nvlink python bench/nvlink_single_proc.py --mib 1024 --objects 16 --iters 1 --algo xxh3 > gpurun_out/r2/ncu/nvlink.log 2>&1
timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_gpu_stack.py -x -q -k "multi or mailbox" > gpurun_out/r2/tests_n2.log 2>&1
tail -3 gpurun_out/r2/tests_n2.log
