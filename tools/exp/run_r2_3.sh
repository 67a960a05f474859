cd "$(dirname "$0")/../.."
python -m pytest tests/test_ops_gpu.py tests/test_gen_layout_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -3
for g in 0 99 0 99; do UMV_RASTER_G=$g python tools/e2e_phases.py /dev/null 2>/dev/null | grep -E "forward_cache_update_vit" | tail -1 | sed "s/^/G=$g /"; done
for g in 0 99; do UMV_RASTER_G=$g python tools/t2i_trace.py gpurun_out/t2i_g$g.md > /dev/null 2>&1; head -1 gpurun_out/t2i_g$g.md | cut -c1-160; done
ONLY="prefill down" UMV_NCU=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_2cta -s 30 -c 18 --csv --log-file gpurun_out/r2_raster_dram.csv python tools/exp/raster_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_raster_dram.csv')))
h=next(i for i,r in enumerate(rows) if 'Kernel Name' in r); H=rows[h]
ki,mi,vi,ii=H.index('Kernel Name'),H.index('Metric Name'),H.index('Metric Value'),H.index('ID')
d={}
for r in rows[h+1:]:
    d.setdefault(r[ii],{})[r[mi]]=r[vi]
for k,v in list(d.items())[:18]: print(k, v)
PY
