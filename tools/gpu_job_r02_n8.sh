set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r02_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --rows 20000000 --steps 2 --no-e2e --no-modes --no-cfd --no-cpu > gpurun_out/bench_r02_n8b.json 2> gpurun_out/bench_r02_n8b.err
tail -3 gpurun_out/r02_pytest_multi.log
