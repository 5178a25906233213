mkdir -p gpurun_out
python -m pytest tests/test_engine_gpu.py tests/test_encoder_gpu.py tests/test_txfm_gpu.py tests/test_misc_gpu.py -x -q -k "engine or encoder or encode_tus or stats" > gpurun_out/t_engine.log 2>&1; tail -15 gpurun_out/t_engine.log
python tools/stats_bench.py > gpurun_out/stats_bench.json 2> gpurun_out/stats_bench.err; cat gpurun_out/stats_bench.json | tail -40
python tools/encode_compare.py --width 1920 --height 1080 --frames 160 --qp 43 --variants ref_simd,cuda_simd,ref_simd,cuda_simd --profile --no-recon > gpurun_out/enc1080_fuse.log 2>&1; tail -30 gpurun_out/enc1080_fuse.log
