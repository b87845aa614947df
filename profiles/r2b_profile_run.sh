# round-2 (second half) evidence: plain runs first (must exit 0), then the ncu passes of the SAME commands
set -x
python profiles/timeline_step.py > gpurun_out/r2b_timeline_step.txt 2>&1 && cp gpurun_out/timeline_step.tsv gpurun_out/r2b_timeline_step.tsv
python profiles/timeline_step.py one_stream > gpurun_out/r2b_timeline_one_stream.txt 2>&1
python profiles/diag_step_branches.py > gpurun_out/r2b_step_branches.json 2> /dev/null
python profiles/diag_cone.py > gpurun_out/r2b_cone.json 2> /dev/null
(cd profiles && python diag_gnn_wgrad.py > ../gpurun_out/r2b_gnn_wgrad.json 2> /dev/null; python diag_gemm_modes.py > ../gpurun_out/r2b_gemm_modes.jsonl 2> /dev/null)
python bench.py --profile-step > gpurun_out/r2b_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_launches_step.csv python bench.py --profile-step > gpurun_out/r2b_ncu_step.log 2>&1
python profiles/dev_gnn_persist.py c2 > gpurun_out/r2b_plain_gnn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gnn_cell_fwd -s 1300 -c 1 -f -o gpurun_out/r2b_prof_gnn_cell_fwd_h16 python profiles/dev_gnn_persist.py c2 > gpurun_out/r2b_ncu_gnn.log 2>&1
tail -3 gpurun_out/r2b_ncu_step.log gpurun_out/r2b_ncu_gnn.log
