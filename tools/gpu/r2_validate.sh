# round-2 validation of the not-yet-measured pieces (run on the GPU box: gpurun -- 'bash tools/gpu/r2_validate.sh')
mkdir -p gpurun_out
python -m pytest tests/test_ois_gpu.py tests/test_pa_gpu.py tests/test_tf_gpu.py tests/test_misc_gpu.py tests/test_txfm_gpu.py tests/test_rtcd_install_gpu.py tests/test_engine_gpu.py tests/test_encoder_gpu.py -q -k "test_ois_gpu or test_pa_gpu or tf or temporal or stats or encode_tus or engine or encoder" > gpurun_out/t_validate.log 2>&1; tail -25 gpurun_out/t_validate.log
python tools/stats_bench.py > gpurun_out/stats_bench.json 2> gpurun_out/stats_bench.err; tail -40 gpurun_out/stats_bench.json
python tools/encode_compare.py --width 1920 --height 1080 --frames 160 --qp 43 --variants ref_simd,cuda_simd,ref_simd,cuda_simd --profile --no-recon > gpurun_out/enc1080_fuse.log 2>&1; tail -30 gpurun_out/enc1080_fuse.log
oracle/_ref/cpu_kernel_bench oracle/_ref/simd/libSvtAv1EncSimd.so > gpurun_out/cpu_kernels_c_vs_simd.json 2>&1; tail -3 gpurun_out/cpu_kernels_c_vs_simd.json
