# the round-2 validation job (one GPU):  gpurun --timeout 1800 -- 'bash scripts/gpu_job_r2.sh'
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
timeout 600 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; tail -c 300 gpurun_out/bench_r2_default.err
