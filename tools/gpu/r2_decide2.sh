mkdir -p gpurun_out
for v in 0 1 0 1; do
  python tools/encode_compare.py --width 1920 --height 1080 --frames 160 --qp 43 --variants cuda_simd --profile --no-recon --env SVT_CUDA_CDEF_DECIDE=$v > gpurun_out/enc_decideab_$v.log 2>&1
  echo "device decision $v: $(grep -o '"fps": [0-9.]*' gpurun_out/enc_decideab_$v.log | tail -1) $(grep -o 'cdef gpu: [0-9]* calls, [0-9.]* ms total wall in stage threads, [0-9.]* ms/call' gpurun_out/enc_decideab_$v.log | tail -1) $(grep -o 'finish_cdef_search [0-9.]*, CDEF engine calls [0-9.]*' gpurun_out/enc_decideab_$v.log | tail -1)"
done
bash tools/gpu/r2_minb.sh
python tools/kernel_bench.py > gpurun_out/kernel_bench_r2f.json 2> gpurun_out/kernel_bench_r2f.err; python -c "
import json; d=json.load(open('gpurun_out/kernel_bench_r2f.json'))
for g in d['geometries']:
    for r in g['rows']:
        if 'decide' in r['entry']: print(g['geometry'], r['entry'][:60], r['ms'])
"
