mkdir -p gpurun_out
STEPS=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel|attn_decode_kernel|add_rmsnorm_splitk" --launch-skip 14 -c 7 -o gpurun_out/r1_decode_full -f python tools/profile_decode.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
STEPS=2 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches2.csv python tools/profile_decode.py > gpurun_out/prof2.log 2>&1
tail -1 gpurun_out/prof2.log
