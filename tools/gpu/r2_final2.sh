# final state of round 2: the whole GPU suite, smoke, the bench lines, then the device-decision A/B and two small experiments
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/t_full_gpu2.log 2>&1; tail -4 gpurun_out/t_full_gpu2.log
python __graft_entry__.py smoke > gpurun_out/smoke3.log 2>&1; tail -1 gpurun_out/smoke3.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2e_ref_n1.json 2> gpurun_out/bench_r2e_ref_n1.err; cut -c1-200 gpurun_out/bench_r2e_ref_n1.json
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2e_n1.json 2> gpurun_out/bench_r2e_n1.err; cut -c1-200 gpurun_out/bench_r2e_n1.json
bash tools/gpu/r2_decide2.sh
