set -x
timeout 300 python -m pytest tests/test_sidecar_gpu.py tests/test_shim_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3
timeout 300 python scripts/bench_concurrent.py --direct 0 --procs 4,16,64 --seconds 3 > gpurun_out/r2_sidecar_callers_l0.json 2> gpurun_out/sidecar.err; grep sidecar gpurun_out/r2_sidecar_callers_l0.json | cut -c1-400 | head -4; tail -c 400 gpurun_out/sidecar.err
timeout 300 python scripts/bench_concurrent.py --direct 0 --procs 16,64 --seconds 3 --linger-us 40 > gpurun_out/r2_sidecar_callers_l40.json 2> gpurun_out/sidecar40.err; grep sidecar gpurun_out/r2_sidecar_callers_l40.json | cut -c1-400 | head -3; tail -c 400 gpurun_out/sidecar40.err
