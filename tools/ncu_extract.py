"""Extract the judged metrics of an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small CSV under profiles/.

usage: python tools/ncu_extract.py gpurun_out/X.ncu-rep profiles/X_raw.csv [kernel-name-substring]
One block per captured launch: kernel name, then metric,unit,value rows for time, launch configuration, DRAM / L2 bytes,
issue and stall metrics (the ones B200_PROFILING.md names).
"""
import csv, subprocess, sys

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__shared_mem_per_block", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput", "dram__cycles_active.avg.pct",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active",
        "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_thread_inst_executed_op_d")
SKIP = (".min", ".max", "pcsamp", "_not_issued")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s (ncu --set full --clock-control none); one block per captured launch\n" % rep.split("/")[-1])
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            if pat and pat not in d.get("Kernel Name", ""):
                continue
            f.write("kernel,,%s\n" % d.get("Kernel Name", "").replace(",", ";"))
            for k, u, v in zip(hdr, units, row):
                if any(s in k for s in KEEP) and not any(s in k for s in SKIP) and v not in ("", "n/a"):
                    f.write("%s,%s,%s\n" % (k, u, v))
            f.write("\n")


if __name__ == "__main__":
    main()
