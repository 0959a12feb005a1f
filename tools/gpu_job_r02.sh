set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cfd.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_pytest_cfd.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_cfd.log
timeout 900 python bench.py --rows 20000000 --steps 2 --no-e2e --no-modes --no-train --no-cpu > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err
tail -5 gpurun_out/r02_pytest_cfd.log
