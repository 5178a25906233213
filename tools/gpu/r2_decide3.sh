mkdir -p gpurun_out
python -m pytest tests/test_cdef_gpu.py tests/test_engine_gpu.py tests/test_encoder_gpu.py -q -k "decide or decision or md5 or conformance" > gpurun_out/t_decide3.log 2>&1; tail -4 gpurun_out/t_decide3.log
python __graft_entry__.py smoke > gpurun_out/smoke4.log 2>&1; tail -1 gpurun_out/smoke4.log | cut -c1-120
python tools/kernel_bench.py > gpurun_out/kernel_bench_r2g.json 2> gpurun_out/kernel_bench_r2g.err; python -c "
import json; d=json.load(open('gpurun_out/kernel_bench_r2g.json'))
for g in d['geometries']:
    for r in g['rows']:
        if 'decide' in r['entry']: print(g['geometry'], r['entry'][:60], r['ms'])
"
python bench.py --config 2160p10 --steps 4 --warmup 1 > gpurun_out/bench_r2e_2160p10.json 2> gpurun_out/bench_r2e_2160p10.err; cut -c1-220 gpurun_out/bench_r2e_2160p10.json
for v in 0 1; do
  python tools/encode_compare.py --width 3840 --height 2160 --bits 10 --preset 6 --frames 48 --qp 43 --variants cuda_simd --profile --no-recon --env SVT_CUDA_CDEF_DECIDE=$v > gpurun_out/enc4k_decide_$v.log 2>&1
  echo "4K device decision $v: $(grep -o '"fps": [0-9.]*' gpurun_out/enc4k_decide_$v.log | tail -1) $(grep -o 'cdef gpu: [0-9]* calls, [0-9.]* ms total wall in stage threads, [0-9.]* ms/call' gpurun_out/enc4k_decide_$v.log | tail -1) $(grep -o 'finish_cdef_search [0-9.]*, CDEF engine calls [0-9.]*' gpurun_out/enc4k_decide_$v.log | tail -1) $(grep -o '"ivf_md5": "[0-9a-f]*"' gpurun_out/enc4k_decide_$v.log | tail -1)"
done
