set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/cfd_bench.py 600000 > gpurun_out/tc5_wide.txt 2>&1
BALER_B200_TC5_NARROW=1 python tools/cfd_bench.py 600000 > gpurun_out/tc5_narrow.txt 2>&1
cat gpurun_out/tc5_wide.txt gpurun_out/tc5_narrow.txt
timeout 900 python -m pytest tests/test_gpu_cfd.py -m gpu -x -q 2>&1 | tail -5
