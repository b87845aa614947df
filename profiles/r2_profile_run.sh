# round-2 evidence: plain runs first (must exit 0), then the ncu passes of the SAME commands
set -x
python bench.py --profile-step > gpurun_out/r2_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --profile-step > gpurun_out/r2_ncu_step.log 2>&1
python profiles/dev_gnn_persist.py c2 > gpurun_out/r2_plain_gnn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gnn_persist_fwd -s 5 -c 1 -f -o gpurun_out/r2_prof_gnn_persist_fwd python profiles/dev_gnn_persist.py c2 > gpurun_out/r2_ncu_gnn.log 2>&1
tail -3 gpurun_out/r2_ncu_step.log gpurun_out/r2_ncu_gnn.log
