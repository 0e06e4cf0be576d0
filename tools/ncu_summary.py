"""Text summary of .ncu-rep captures (ncu --set full) for profiles/:  python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "smsp__warps_eligible.avg.per_cycle_active"]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            print("## %s   [%s]" % (d["Kernel Name"], rep.split("/")[-1]))
            for k in KEYS:
                if k in d:
                    print("%-70s %s %s" % (k, d[k], u[k]))
            st = [(float(d[k]) if d[k] not in ("", "n/a") else 0.0,
                   k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                  for k in hdr if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k]
            print("warp stall reasons (warps per issue-active cycle):")
            for v, k in sorted(st, reverse=True)[:9]:
                print("    %.2f %s" % (v, k))
            print()


if __name__ == "__main__":
    main()
