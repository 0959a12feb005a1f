set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:tc_train_kernel -c 1 -f -o gpurun_out/r02_train_tc_dbn python tools/train_prof_dbn.py dbn > gpurun_out/r02_ncu_dbn.log 2>&1
timeout 600 $NCU -k regex:gemm_tc5 -s 5 -c 10 -f -o gpurun_out/r02_gemm_tc5 python tools/cfd_bench.py 75776 > gpurun_out/r02_ncu_tc5.log 2>&1
ls -la gpurun_out/r02_train_tc_dbn.ncu-rep gpurun_out/r02_gemm_tc5.ncu-rep
