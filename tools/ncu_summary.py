"""Summarise an .ncu-rep (ncu --set full) into markdown: headline metrics, stall reasons, DRAM traffic.
Usage: python tools/ncu_summary.py report.ncu-rep [units_per_launch unit_name algorithmic_bytes_per_unit]"""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (tcgen05) active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__sass_inst_executed_op_local_ld.sum", "local loads"), ("sm__sass_inst_executed_op_local_st.sum", "local stores"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        col = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        print("### `%s`\n" % col["Kernel Name"][0])
        print("| metric | value |\n|---|---|")
        for k, name in KEYS:
            if k in col and col[k][0] not in ("", "n/a"):
                print("| %s (`%s`) | %s %s |" % (name, k, col[k][0], col[k][1]))
        rd = wr = None
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(col["dram__bytes_read.sum"][0]) * scale[col["dram__bytes_read.sum"][1]]
            wr = float(col["dram__bytes_write.sum"][0]) * scale[col["dram__bytes_write.sum"][1]]
            us = float(col["gpu__time_duration.sum"][0]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}[col["gpu__time_duration.sum"][1]]
            print("| DRAM traffic (read + write) | %.4f GB -> %.0f GB/s under ncu |" % ((rd + wr) / 1e9, (rd + wr) / us / 1e3))
            if len(sys.argv) >= 5:
                n, unit, b = float(sys.argv[2]), sys.argv[3], float(sys.argv[4])
                print("| algorithmic bytes | %.4f GB (%g %s x %g B) -> traffic / algorithmic = %.3f |" % (
                    n * b / 1e9, n, unit, b, (rd + wr) / (n * b)))
        except Exception as e:  # noqa: BLE001
            print("| traffic | n/a (%r) |" % (e,))
        print("\nWarp stall reasons (warps per issue cycle, `smsp__average_warps_issue_stalled_*_per_issue_active`):\n")
        st = []
        for h, (v, _) in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(v), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print(", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
        print()


if __name__ == "__main__":
    main()
