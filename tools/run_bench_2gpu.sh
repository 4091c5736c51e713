python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 24 --warmup 3 > gpurun_out/r1_bench_final3_2gpu.json 2> gpurun_out/bench2.err
tail -c 700 gpurun_out/r1_bench_final3_2gpu.json | head -c 300; echo; cut -c1-260 gpurun_out/r1_bench_final3_2gpu.json | tail -n 1; tail -n 2 gpurun_out/bench2.err
