for g in 32 64 128; do
echo "== L2 fetch granularity $g"
SWEEP_L2_GRAN=$g python scripts/sweep_search.py "0,64,75776,10000,0,2,3200" 2>&1 | grep -E "lpw|granularity"
done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,smsp__thread_inst_executed_per_inst_executed.ratio,lts__d_sectors_fill_device.sum
timeout 900 ncu --replay-mode application --metrics $M --clock-control none -k regex:search_compact_kernel -c 1 --csv --log-file gpurun_out/r73_search_compact_v9_fullsize.csv python bench.py --steps 1 --warmup 0 --cpu-repeats 1 > gpurun_out/r73_ncu.log 2>&1
echo done
