mkdir -p gpurun_out
SVT_B200_STATS_MINB=2 python -m pytest tests/test_misc_gpu.py -q -k "stats" > gpurun_out/t_minb.log 2>&1; tail -3 gpurun_out/t_minb.log
for m in 1 2 1 2; do SVT_B200_STATS_MINB=$m python tools/stats_bench.py > gpurun_out/stats_bench_minb$m.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/stats_bench_minb$m.json')); print('minb $m', [r['ms'] for r in list(d.values())[0]], d['same_results'])"; done
