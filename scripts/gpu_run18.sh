mkdir -p gpurun_out/r2
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
for mode_skip in 3 16 29; do
  echo "=== config 5 (Zipf 1M x 1M x 1K, 200M nnz, R=64), launch skip $mode_skip"
  timeout 600 ncu --metrics $M --clock-control none -k regex:mttkrp_stream -s $mode_skip -c 1 python scripts/quick_bench.py 1000000 200000000 64 3 0 zipf 2>&1 | grep -E "gpu__time|hit_rate|xbar2l1tex|dram__|lts2xbar|data_pipe|issue_active|registers|warps_active|^mode"
done
