cd "$(dirname "$0")/../.."
for tk in 0 128 0 128; do UMV_ATTN_TK=$tk python tools/e2e_phases.py /dev/null 2>/dev/null | grep -E "forward_cache_update_vit" | tail -1 | sed "s/^/TK=$tk /"; done
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_r2i.log 2>&1; tail -2 gpurun_out/pytest_r2i.log
python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2e.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['step_roofline']['frac'], d['t2i']['value'], d['t2i']['roofline']['frac'], d['interleaved']['value'], d['report_gen']['value'])"
