set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cfd.py -m gpu -x -q 2>&1 | tail -5
python tools/cfd_bench.py 600000 > gpurun_out/tc5_coalesced.txt 2>&1
python tools/cfd_bench.py 600000 >> gpurun_out/tc5_coalesced.txt 2>&1
cat gpurun_out/tc5_coalesced.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_cfd_launches2.csv python tools/cfd_bench.py 75776 > /dev/null 2>&1
