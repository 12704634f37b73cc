"""Prints the handful of ncu raw-page metrics the roofline discussion uses.  usage: ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed_op_tma_ld.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("metric,unit," + ",".join("launch%d" % i for i in range(len(data))))
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print("%s,%s,%s" % (h, units[i], ",".join('"%s"' % d[i] for d in data)))


if __name__ == "__main__":
    main(sys.argv[1])
