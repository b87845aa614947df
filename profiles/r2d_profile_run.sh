# last evidence of round 2: ncu --set full of the per-level backward cell kernel (fp16 split) and of the generic GEMM
# after the row-direct epilogue, plain runs first
set -x
python profiles/dev_gnn_persist.py c2 > gpurun_out/r2d_plain_gnn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gnn_cell_bwd -s 1300 -c 1 -f -o gpurun_out/r2d_prof_gnn_cell_bwd_h16 python profiles/dev_gnn_persist.py c2 > gpurun_out/r2d_ncu_gnn.log 2>&1
python profiles/trace_gemm.py > gpurun_out/r2d_plain_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tf_gemm -c 1 -f -o gpurun_out/r2d_prof_tf_gemm python profiles/trace_gemm.py > gpurun_out/r2d_ncu_gemm.log 2>&1
ls -la gpurun_out | grep r2d
