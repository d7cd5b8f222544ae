tag=r2an
export VB_OPTIONS="tc_drain=8"
B="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_knn2_tc4 -s 4 -c 1 -o gpurun_out/prof_tc4d8_$tag -f $B > gpurun_out/ncu_tc4d8_$tag.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_tc4d8_$tag.ncu-rep > gpurun_out/${tag}_ncu_tc4d8.txt 2>&1
grep -E "time_duration|registers|inst_executed.sum|pipe_alu_cycles|pipe_tensor_cycles|issue_active|stalled_(wait|long|short|barrier|branch|math|not_sel|no_inst)|local" gpurun_out/${tag}_ncu_tc4d8.txt
