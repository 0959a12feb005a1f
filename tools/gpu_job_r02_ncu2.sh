set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 10 -c 5 -f -o gpurun_out/r02_gemm_tc5 python tools/cfd_bench.py 75776 > gpurun_out/r02_ncu_tc5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_cfd_launches.csv python tools/cfd_bench.py 75776 > /dev/null 2>&1
timeout 300 python tools/cfd_bench2.py > gpurun_out/r02_cfd_bench2.txt 2>&1
ls -la gpurun_out/r02_gemm_tc5.ncu-rep
