set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --gpu-rotate 3 > gpurun_out/bench_r2_n8_rot3.json 2> gpurun_out/bench_r2_n8_rot3.err; tail -c 600 gpurun_out/bench_r2_n8_rot3.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_r2_n8_rot3.json').read().strip().splitlines()[-1])
    print("N=8 value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "clocks", j.get("clocks"))
    for r in j["config"]["per_rank"]: print("  rank", r["rank"], "gpu", r["gpu"], round(r["ms_per_step_device"],3), round(r["ms_per_step_e2e"],3), r["sm_mhz"], r["stage_ms_per_step"], r["exact_path_queries_per_step"])
    for s in j["config"]["secondary"]:
        print("   ", s.get("name","")[:40], s.get("seconds"), s.get("equals_reference_on_sample",{}).get("ok"), s.get("error"))
except Exception as ex:
    print("no N=8 result", ex)
PY
