python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_r2f.log 2>&1; echo rc=$? >> gpurun_out/pytest_r2f.log; tail -15 gpurun_out/pytest_r2f.log
python tools/t2i_trace.py gpurun_out/r2_t2i_trace_seg.md > /dev/null 2>&1; UMV_GEN_SEG=0 python tools/t2i_trace.py gpurun_out/r2_t2i_trace_packed.md > /dev/null 2>&1
python tools/kernel_vs_library.py gpurun_out/kernel_vs_library.md > gpurun_out/kvl.log 2>&1; tail -3 gpurun_out/kvl.log
