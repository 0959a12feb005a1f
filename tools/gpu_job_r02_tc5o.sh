set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 240 python -m pytest tests/test_gpu_cfd.py -m gpu -x -q 2>&1 | tail -5
timeout -k 10 120 python tools/cfd_bench.py 600000 > gpurun_out/tc5_coalesced.txt 2>&1
BALER_B200_TC5_STREAM=1 timeout -k 10 120 python tools/cfd_bench.py 600000 >> gpurun_out/tc5_coalesced.txt 2>&1
cat gpurun_out/tc5_coalesced.txt
timeout -k 10 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_cfd_launches2.csv python tools/cfd_bench.py 75776 > /dev/null 2>&1
