M=l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for m in 0 1 2; do
VARNET_B200_TC64_GW=$m AB_NTF=37888 timeout 300 ncu --metrics $M --clock-control none -k regex:tc64_var_kernel -s 4 -c 1 --csv --log-file gpurun_out/r2bd_gw$m.csv python scripts/ab_tc64.py --child m$m > gpurun_out/r2bd_gw$m.log 2>&1
done
