set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --rows 20000000 --steps 3 --no-e2e --no-modes --no-cfd > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err
tail -5 gpurun_out/r02_pytest_gpu.log
