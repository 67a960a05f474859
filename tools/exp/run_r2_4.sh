cd "$(dirname "$0")/../.."
python -m pytest tests/test_ops_gpu.py tests/test_vae_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -3
ONLY="prefill q" python tools/exp/raster_probe.py 2>&1 | tail -2
for g in 0 99999 0 99999; do UMV_RASTER_G=$g python tools/e2e_phases.py /dev/null 2>/dev/null | grep -E "forward_cache_update_vit" | tail -1 | sed "s/^/G=$g /"; done
