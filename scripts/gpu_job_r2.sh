set -x
timeout 300 python -m pytest tests/test_encode_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -c 300 gpurun_out/ncu_bench.log
wc -l gpurun_out/r2_launch_list.csv
which nvidia-cuda-mps-control nvidia-smi
if which nvidia-cuda-mps-control; then
  export CUDA_MPS_PIPE_DIRECTORY=/tmp/nvidia-mps CUDA_MPS_LOG_DIRECTORY=/tmp/nvidia-mps-log
  mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
  nvidia-cuda-mps-control -d; sleep 1
  timeout 400 python scripts/bench_concurrent.py --procs 1,4,16,64 --seconds 3 > gpurun_out/r2_concurrent_callers_mps.json 2> gpurun_out/conc_mps.err; tail -c 1200 gpurun_out/r2_concurrent_callers_mps.json; tail -c 500 gpurun_out/conc_mps.err
  echo quit | timeout 20 nvidia-cuda-mps-control
  tail -5 /tmp/nvidia-mps-log/control.log
fi
