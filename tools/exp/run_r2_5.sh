cd "$(dirname "$0")/../.."
for tk in 64 128; do
  echo "== UMV_ATTN_TK=$tk"
  UMV_ATTN_TK=$tk python -m pytest tests/test_ops_gpu.py tests/test_attention_block_gpu.py tests/test_model_gpu.py tests/test_flow_gpu.py tests/test_gen_layout_gpu.py -m gpu -q --timeout 900 2>&1 | tail -3
  UMV_ATTN_TK=$tk python tools/kernel_vs_library.py gpurun_out/kvl_tk$tk.md > /dev/null 2>&1
  grep -A 9 "^## attention" gpurun_out/kvl_tk$tk.md | cut -d'|' -f2,7,8,11,13 
done
