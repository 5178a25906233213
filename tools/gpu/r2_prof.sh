mkdir -p gpurun_out
SVT_B200_ENGINE_TRACE=0 python tools/encode_compare.py --width 1920 --height 1080 --frames 160 --qp 43 --variants cuda_simd,cuda_simd --profile --no-recon > gpurun_out/enc_prof_host.log 2>&1
grep -o 'SVT \[CUDA profile\][^"]*' gpurun_out/enc_prof_host.log | cut -c1-400; grep -o '"fps": [0-9.]*' gpurun_out/enc_prof_host.log
