# final round-2 measurement: the whole GPU suite, then the bench lines (both arms), then configs[2]
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/t_full_gpu.log 2>&1; tail -5 gpurun_out/t_full_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2d_ref_n1.json 2> gpurun_out/bench_r2d_ref_n1.err; cat gpurun_out/bench_r2d_ref_n1.json | cut -c1-400
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2d_n1.json 2> gpurun_out/bench_r2d_n1.err; cat gpurun_out/bench_r2d_n1.json | cut -c1-900
python bench.py --config 2160p10 --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r2d_2160p10_ref.json 2> gpurun_out/bench_r2d_2160p10_ref.err; cat gpurun_out/bench_r2d_2160p10_ref.json | cut -c1-300
python bench.py --config 2160p10 --steps 4 --warmup 1 > gpurun_out/bench_r2d_2160p10.json 2> gpurun_out/bench_r2d_2160p10.err; cat gpurun_out/bench_r2d_2160p10.json | cut -c1-600
