set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches.csv python bench.py --profile-step > gpurun_out/final_ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gnn_.*_fwd --csv --log-file gpurun_out/final_traffic.csv python bench.py --profile-step > gpurun_out/final_ncu_traffic.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gnn_cell_fwd --launch-skip 20 -c 1 -o gpurun_out/final_prof_cell_fwd python bench.py --profile-step > gpurun_out/final_ncu_cell.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:tf_gemm --launch-skip 20 -c 3 -o gpurun_out/final_prof_tf_gemm python bench.py --profile-step > gpurun_out/final_ncu_gemm.log 2>&1
tail -2 gpurun_out/final_gpu_tests.log; cat gpurun_out/final_smoke.log | tail -2
