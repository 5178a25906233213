mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 --steps 12 --warmup 3 > gpurun_out/bench_r2d_ref_n2.json 2> gpurun_out/bench_r2d_ref_n2.err; cut -c1-300 gpurun_out/bench_r2d_ref_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/bench_r2d_n2.json 2> gpurun_out/bench_r2d_n2.err; cut -c1-600 gpurun_out/bench_r2d_n2.json
