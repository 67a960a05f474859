mkdir -p gpurun_out
export UMV_NEAR_ATTN_STAGES=10
for c in "UMV_PF_WO=1 UMV_PF_GU_KB=0" "UMV_PF_WO=1 UMV_PF_GU_KB=20" "UMV_PF_WO=0 UMV_PF_GU_KB=0 UMV_PF_N2_KB=20"; do echo "== $c"; env $c timeout 200 python tools/decode_trace.py 2>&1 | grep -A12 "per kernel class" | tail -8; done
