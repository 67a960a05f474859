mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gemm_2cta_kernel|attn_tc_kernel" --launch-skip 120 -c 6 -o gpurun_out/r1_prefill_tc -f python tools/prefill_trace.py > gpurun_out/ncu_tc.log 2>&1
tail -1 gpurun_out/ncu_tc.log
timeout 300 python tools/t2i_trace.py gpurun_out/t2i_trace_final.md > /dev/null 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-400
