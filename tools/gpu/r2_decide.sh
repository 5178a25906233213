mkdir -p gpurun_out
python -m pytest tests/test_cdef_gpu.py tests/test_engine_gpu.py tests/test_encoder_gpu.py -q -k "decide or decision or encoder" > gpurun_out/t_decide.log 2>&1; tail -12 gpurun_out/t_decide.log
python __graft_entry__.py smoke > gpurun_out/smoke2.log 2>&1; tail -2 gpurun_out/smoke2.log
for v in 1 0 1 0; do
  SVT_CUDA_CDEF_DECIDE=$v python tools/encode_compare.py --width 1920 --height 1080 --frames 160 --qp 43 --variants cuda_simd --profile --no-recon > gpurun_out/enc_decide_$v.log 2>&1
  echo "device decision $v: $(grep -o '"fps": [0-9.]*' gpurun_out/enc_decide_$v.log | tail -1) $(grep -o 'cdef gpu: [0-9]* calls, [0-9.]* ms total wall in stage threads, [0-9.]* ms/call' gpurun_out/enc_decide_$v.log | tail -1) $(grep -o 'finish_cdef_search [0-9.]*, CDEF engine calls [0-9.]*' gpurun_out/enc_decide_$v.log | tail -1) $(grep -o '"ivf_md5": "[0-9a-f]*"' gpurun_out/enc_decide_$v.log | tail -1)"
done
