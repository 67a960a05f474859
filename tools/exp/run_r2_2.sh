python -m pytest tests/test_fullsize_gpu.py tests/test_model_gpu.py tests/test_scheduler_gpu.py tests/test_attention_block_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -5
python tools/decode_time.py 2>/dev/null | tail -2
UMV_TAIL_NORM=0 python tools/decode_time.py 2>/dev/null | tail -1
python tools/decode_time.py 2>/dev/null | tail -1
UMV_TAIL_NORM=0 python tools/decode_time.py 2>/dev/null | tail -1
