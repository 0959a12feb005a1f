set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_cfd.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3
timeout -k 10 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 5 -c 10 -f -o gpurun_out/r02_gemm_tc5 python tools/cfd_bench.py 151552 > gpurun_out/r02_ncu_tc5.log 2>&1
timeout -k 10 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_cfd_launches2.csv python tools/cfd_bench.py 151552 > /dev/null 2>&1
