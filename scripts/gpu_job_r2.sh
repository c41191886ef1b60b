timeout 60 scripts/umma_probe 200 5000 300 5 | tail -3
timeout 60 scripts/umma_probe 300 70000 300 10 | tail -3
timeout 60 scripts/umma_probe 1000 3000000 300 1 perf | grep -E "rep 4|PROBE|candidates"
timeout 300 python scripts/bench_prefilter.py 2>&1 | tail -12
timeout 300 python -m pytest tests/test_prefilter_gpu.py tests/test_rerank_gpu.py tests/test_vector_udfs.py tests/test_append_gpu.py -q --timeout 280 2>&1 | tail -3
