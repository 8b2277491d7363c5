python scripts/backup_bench.py 75776 4 2>&1 | tail -3
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,launch__registers_per_thread,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
timeout 600 ncu --metrics $M --clock-control none -k regex:"backup_uct_kernel|backup_hypothesis_kernel|root_merge_kernel|root_finish_kernel" --csv --log-file gpurun_out/a11_backup_kernels.csv python scripts/backup_bench.py 75776 4 > gpurun_out/a11_backup_ncu.log 2>&1
tail -2 gpurun_out/a11_backup_ncu.log
python scripts/backup_bench.py 1000000 6 2>&1 | tail -3
timeout 600 ncu --metrics $M --clock-control none -k regex:"backup_uct_kernel|backup_hypothesis_kernel|root_merge_kernel|root_finish_kernel" --csv --log-file gpurun_out/a11_backup_kernels_1m.csv python scripts/backup_bench.py 1000000 6 > gpurun_out/a11_backup_ncu_1m.log 2>&1
