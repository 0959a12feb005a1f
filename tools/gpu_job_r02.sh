set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu_final.log
timeout 1500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
python tools/train_prof_dbn.py dbn > gpurun_out/r02_prof_dbn.txt 2>&1
tail -3 gpurun_out/r02_pytest_gpu_final.log; tail -2 gpurun_out/r02_smoke.log
