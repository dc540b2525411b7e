#!/usr/bin/env python
"""One line per captured launch of an .ncu-rep (`ncu --set full`): duration, DRAM bytes, tensor-pipe activity, L2 traffic.
   python tools/ncu_table.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys


def col(idx, row, *names):
    for n in names:
        if n in idx and row[idx[n]] not in ("", "n/a"):
            return float(row[idx[n]].replace(",", ""))
    return float("nan")


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    unit = lambda n: units[idx[n]] if n in idx else ""
    scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
    tscale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
    print(f"{'#':>2} {'kernel':<34} {'ms':>8} {'dram rd GB':>10} {'dram wr GB':>10} {'dram %':>7} {'tensor act %':>12} {'L2->SM GB':>10} "
          f"{'L2 hit %':>8} {'SM GHz':>7} {'regs':>5} {'block':>6} {'smem KB':>8}")
    for i, r in enumerate(rows[2:], 1):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[-34:]
        ms = col(idx, r, "gpu__time_duration.sum") * tscale.get(unit("gpu__time_duration.sum"), 1e-6)
        rd = col(idx, r, "dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1e-9)
        wr = col(idx, r, "dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1e-9)
        l2 = col(idx, r, "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum") * scale.get(unit("l1tex__m_xbar2l1tex_read_bytes.sum") or unit("lts__t_bytes.sum"), 1e-9)
        ghz = col(idx, r, "sm__cycles_elapsed.avg.per_second")
        ghz *= {"hz": 1e-9, "Hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0, "cycle/nsecond": 1.0, "cycle/usecond": 1e-3, "cycle/second": 1e-9}.get(unit("sm__cycles_elapsed.avg.per_second"), 1.0)
        smem = col(idx, r, "launch__shared_mem_per_block_dynamic") * {"byte/block": 1e-3, "Kbyte/block": 1.0}.get(unit("launch__shared_mem_per_block_dynamic"), 1e-3)
        print(f"{i:>2} {name:<34} {ms:8.3f} {rd:10.2f} {wr:10.2f} {col(idx, r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f} "
              f"{col(idx, r, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):12.1f} "
              f"{l2:10.1f} {col(idx, r, 'lts__t_sector_hit_rate.pct'):8.1f} {ghz:7.2f} {col(idx, r, 'launch__registers_per_thread'):5.0f} "
              f"{col(idx, r, 'launch__block_size'):6.0f} {smem:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
