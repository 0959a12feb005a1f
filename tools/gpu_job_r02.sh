set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu3.log
timeout 1500 python bench.py > gpurun_out/bench_r02_e.json 2> gpurun_out/bench_r02_e.err
tail -4 gpurun_out/r02_pytest_gpu3.log
