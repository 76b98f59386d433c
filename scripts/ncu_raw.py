"""Print the key raw metrics of an .ncu-rep (first profiled kernel)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; vals = rows[2] if len(rows) > 2 else rows[1]
want = ['Kernel Name','gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__occupancy_limit_warps','launch__grid_size','launch__block_size','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_fmaheavy','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_lsu.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','launch__shared_mem_dynamic','sm__inst_executed_pipe_uniform.sum','sm__inst_executed_pipe_adu.sum','sm__inst_executed_pipe_cbu.sum','smsp__average_warp','sm__throughput','launch__waves','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','sm__pipe']
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want) and ('pct_of_peak_sustained_elapsed' not in h or 'pipe' in h):
        print(h, '=', vals[i], rows[1][i])
