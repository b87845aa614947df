# end-of-round-2 evidence: plain runs first (must exit 0), then the ncu passes of the SAME commands
set -x
python profiles/timeline_step.py > gpurun_out/r2c_timeline_step.txt 2>&1 && cp gpurun_out/timeline_step.tsv gpurun_out/r2c_timeline_step.tsv
python profiles/diag_step_branches.py > gpurun_out/r2c_step_branches.json 2> /dev/null
(cd profiles && python diag_gnn_wgrad.py > ../gpurun_out/r2c_gnn_wgrad.json 2> /dev/null)
python profiles/time_selfmlp.py > gpurun_out/r2c_time_selfmlp.txt 2>&1
python bench.py --profile-step > gpurun_out/r2c_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches_step.csv python bench.py --profile-step > gpurun_out/r2c_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:selfmlp_gen_fwd -s 3 -c 1 -f -o gpurun_out/r2c_prof_selfmlp_fwd python profiles/time_selfmlp.py > gpurun_out/r2c_ncu_selfmlp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:selfmlp_gen_wgrad2 -s 3 -c 1 -f -o gpurun_out/r2c_prof_selfmlp_wgrad2 python profiles/time_selfmlp.py >> gpurun_out/r2c_ncu_selfmlp.log 2>&1
ls -la gpurun_out | grep r2c
