mkdir -p gpurun_out
python -m pytest tests/test_misc_gpu.py -q -k "stats" > gpurun_out/t_validate3.log 2>&1; tail -4 gpurun_out/t_validate3.log
python tools/stats_bench.py > gpurun_out/stats_bench2.json 2> gpurun_out/stats_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/stats_bench2.json')); print([r['ms'] for r in list(d.values())[0]], d['speedup'], d['same_results'])"
python tools/encode_compare.py --width 1920 --height 1080 --frames 160 --qp 43 --variants ref_simd,cuda_simd:me,cuda_simd:me+cdef,cuda_simd,ref_simd,cuda_simd:me,cuda_simd:me+cdef,cuda_simd --profile --no-recon > gpurun_out/enc1080_stages.log 2>&1; grep -o '"variant": "[^"]*", "rc": [0-9]*, "fps": [0-9.]*' gpurun_out/enc1080_stages.log
