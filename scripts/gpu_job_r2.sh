set -x
timeout 600 python -m pytest tests/test_rerank_gpu.py tests/test_shim_gpu.py tests/test_sidecar_gpu.py -m gpu -q --timeout 300 2>&1 | tail -4
