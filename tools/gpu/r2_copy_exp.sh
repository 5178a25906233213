# host-copy threading experiment (engine staging copies) + the kernel_bench rows of the new entries
mkdir -p gpurun_out
python tools/kernel_bench.py > gpurun_out/kernel_bench_r2e.json 2> gpurun_out/kernel_bench_r2e.err; python -c "
import json; d=json.load(open('gpurun_out/kernel_bench_r2e.json'))
for g in d['geometries']:
    print(g['geometry'], g['same_size_copy'])
    for r in g['rows']: print('  %-90s %.4f ms  %.0f GB/s  %.3f of peak' % (r['entry'][:90], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak']))
"; tail -3 gpurun_out/kernel_bench_r2e.err
for cfg in "4096 4" "512 4" "512 8" "4096 4" "512 4"; do set -- $cfg
  SVT_B200_COPY_MIN_KB=$1 SVT_B200_COPY_THREADS=$2 python tools/encode_compare.py --width 1920 --height 1080 --frames 160 --qp 43 --variants cuda_simd --profile --no-recon > gpurun_out/enc_copy_$1_$2.log 2>&1
  echo "min_kb $1 threads $2: $(grep -o '"fps": [0-9.]*' gpurun_out/enc_copy_$1_$2.log | tail -1) $(grep -o 'host staging copies [0-9.]*' gpurun_out/enc_copy_$1_$2.log | tail -1) $(grep -o 'cdef gpu: [0-9]* calls, [0-9.]* ms total wall in stage threads, [0-9.]* ms/call' gpurun_out/enc_copy_$1_$2.log | tail -1)"
done
