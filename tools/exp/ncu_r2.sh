#!/bin/bash
# Round-2 ncu artefacts: decode launch list + full capture of the decode kernels; flow-step launch list; attention + pair-linear full capture.
cd "$(dirname "$0")/../.."
export UMV_GRAPH=0
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_decode.csv python tools/profile_decode.py > gpurun_out/ncu_r2_a.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_decode_kernel|add_rmsnorm_splitk" -c 14 -f -o gpurun_out/r2_decode_full python tools/profile_decode.py > gpurun_out/ncu_r2_b.log 2>&1
UMV_NCU=1 B=4 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_t2i.csv python tools/t2i_trace.py /dev/null > gpurun_out/ncu_r2_c.log 2>&1
UMV_NCU=1 B=4 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attn_tc_kernel|gemm_2cta_kernel" -s 40 -c 6 -f -o gpurun_out/r2_t2i_full python tools/t2i_trace.py /dev/null > gpurun_out/ncu_r2_d.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep gpurun_out/r2_launches_*.csv; tail -2 gpurun_out/ncu_r2_*.log
