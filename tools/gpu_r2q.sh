# round 2, call Q (GPU box): packed-fp32 build - times, then ncu of k_caves / k_fill_rock for the base and the packed library
OUT=gpurun_out/r2q; mkdir -p $OUT
for v in base q9r8; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 2>&1 | tail -1; done | tee $OUT/variants.txt
python tools/variant_time.py 128 2>&1 | tail -1 | tee -a $OUT/variants.txt
M=smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,gpu__time_duration.sum,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
for v in base mmgen; do
  L=$PWD/mega-minecraft_b200/lib$v.so; [ $v = base ] && L=$PWD/mega-minecraft_b200/libmmgen_base.so
  for K in k_caves:2 k_fill_rock:9; do
    MMGEN_LIB=$L timeout 600 ncu --metrics $M --clock-control none -k regex:${K%%:*} -s ${K##*:} -c 1 --csv --log-file $OUT/${K%%:*}_$v.csv python tools/profile_driver.py 128 1 > $OUT/ncu_${K%%:*}_$v.log 2>&1
  done
done
ls $OUT
