set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:tc_train_kernel -c 1 -f -o gpurun_out/r02_train_tc_ae python tools/train_prof_dbn.py ae > gpurun_out/r02_ncu_ae.log 2>&1
timeout 600 $NCU -k regex:tc_train_kernel -c 1 -f -o gpurun_out/r02_train_tc_dbn python tools/train_prof_dbn.py dbn > gpurun_out/r02_ncu_dbn.log 2>&1
timeout 600 $NCU -k regex:dense_layer_tc -s 3 -c 2 -f -o gpurun_out/r02_dense_layer_tc python tools/cfd_bench.py 65536 > gpurun_out/r02_ncu_cfd.log 2>&1
timeout 600 $NCU -k regex:chain_tc4 -s 6 -c 2 -f -o gpurun_out/r02_chain_tc4 python bench.py --rows 20000000 --steps 1 --no-e2e --no-cpu --no-train --no-cfd --no-modes > gpurun_out/r02_ncu_tc4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --rows 20000000 --steps 2 --warmup 3 --no-e2e --no-cpu --no-modes --no-cfd > gpurun_out/r02_launches_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
