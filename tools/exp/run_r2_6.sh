cd "$(dirname "$0")/../.."
for tk in 0 128 0 128; do UMV_ATTN_TK=$tk python tools/e2e_phases.py /dev/null 2>/dev/null | grep -E "forward_cache_update_vit" | tail -1 | sed "s/^/TK=$tk /"; done
for tk in 0 128; do UMV_ATTN_TK=$tk python tools/t2i_trace.py gpurun_out/t2i_tk$tk.md > /dev/null 2>&1; head -1 gpurun_out/t2i_tk$tk.md | cut -c1-160; grep "attn_tc" gpurun_out/t2i_tk$tk.md; done
UMV_ATTN_TK=0 python tools/prefill_trace.py gpurun_out/prefill_tk0.md > /dev/null 2>&1; grep "attn_tc" gpurun_out/prefill_tk0.md | head -8
