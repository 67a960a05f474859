mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_fwd_kernel|rope_append_kernel|add_rmsnorm_kernel" --launch-skip 600 -c 4 -o gpurun_out/r1_t2i_attn -f python tools/t2i_trace.py > gpurun_out/ncu_t2i.log 2>&1
tail -2 gpurun_out/ncu_t2i.log
