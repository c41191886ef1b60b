set -x
timeout 900 python -m pytest tests/test_knn_join_gpu.py tests/test_full_size_gpu.py tests/test_build_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/gputests.log 2>&1; tail -4 gpurun_out/gputests.log
timeout 600 python bench.py --secondary 4 --no-cpu-baseline --steps 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; tail -c 300 gpurun_out/bench_tmp.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_tmp.json').read().strip().splitlines()[-1])
for s in j["config"]["secondary"]:
    print("   ", s.get("name","")[:40], s.get("seconds"), s.get("stage_ms_rank0"), s.get("equals_reference_on_sample",{}).get("ok"), s.get("error"))
PY
