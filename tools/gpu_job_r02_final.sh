set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu_final.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
for c in 75776 151552; do BALER_B200_LAYER_CHUNK=$c timeout -k 10 120 python tools/cfd_bench.py 600000 2>&1 | grep auto | sed "s/^/chunk $c: /" >> gpurun_out/r02_cfd_chunk.txt; done
tail -3 gpurun_out/r02_pytest_gpu_final.log; tail -2 gpurun_out/r02_smoke.log; cat gpurun_out/r02_cfd_chunk.txt
