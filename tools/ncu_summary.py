"""Summarise an .ncu-rep (read on the CPU box): one line per profiled launch with the metrics the roofline needs."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"),
        ("sm__cycles_elapsed.avg", "cycles")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        name = d[idx["Kernel Name"]]
        short = name.split("(")[0].split("::")[-1][:28]
        parts = []
        for key, label in WANT:
            if key in idx:
                v = d[idx[key]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                parts.append(f"{label}={v}{units[idx[key]] if label in ('time', 'dram_rd', 'dram_wr', 'l2->sm') else ''}")
        print(f"{d[idx['ID']]:>3} {short:28s} grid={d[idx['Grid Size']]:>14s} " + " ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
