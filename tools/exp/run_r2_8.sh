cd "$(dirname "$0")/../.."
python -m pytest tests/test_model_gpu.py tests/test_scheduler_gpu.py tests/test_e2e_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -3
python tools/e2e_phases.py gpurun_out/e2e_fused.md > /dev/null 2>&1; grep -A8 "synchronised" gpurun_out/e2e_fused.md
python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2f.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['step_roofline']['frac'], d['t2i']['value'], d['t2i']['roofline']['frac'], d['interleaved']['value'], d['report_gen']['value'])"
