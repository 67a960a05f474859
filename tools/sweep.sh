mkdir -p gpurun_out
STEPS=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel|attn_decode_kernel|add_rmsnorm_splitk" --launch-skip 14 -c 7 -o gpurun_out/r1_decode_full2 -f python tools/profile_decode.py > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_tc_kernel|attn_fwd_kernel" --launch-skip 40 -c 3 -o gpurun_out/r1_prefill_attn -f python tools/prefill_trace.py > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log
STEPS=2 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches3.csv python tools/profile_decode.py > gpurun_out/prof3.log 2>&1
timeout 200 python tools/decode_trace.py gpurun_out/trace_final.md > /dev/null 2>&1
timeout 300 python tools/t2i_trace.py gpurun_out/t2i_trace_final.md > /dev/null 2>&1
ls gpurun_out | tail -5
