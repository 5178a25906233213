# second validation pass: open-loop intra search + an ncu capture of the tensor-core Wiener statistics kernel
mkdir -p gpurun_out
python -m pytest tests/test_ois_gpu.py tests/test_encoder_gpu.py -q -k "test_ois_gpu or parity_switches" > gpurun_out/t_validate2.log 2>&1; tail -8 gpurun_out/t_validate2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stats_mma_kernel -c 1 -o gpurun_out/stats_mma python tools/stats_bench.py --child > gpurun_out/ncu_stats_mma.log 2>&1; tail -3 gpurun_out/ncu_stats_mma.log
SVT_B200_STATS_IMAD=1 timeout 600 ncu --set full --clock-control none -k regex:stats_kernel -c 1 -o gpurun_out/stats_imad python tools/stats_bench.py --child > gpurun_out/ncu_stats_imad.log 2>&1; tail -3 gpurun_out/ncu_stats_imad.log
ls -la gpurun_out/*.ncu-rep
