"""Summarise an .ncu-rep (ncu --set full) into a small CSV that can be committed under profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/prof_summary.csv
One row per profiled launch with the counters DESIGN.md / bench.py quote (duration, DRAM bytes, tensor / XU pipe
utilisation, issue utilisation, occupancy, registers, grid)."""
import csv
import subprocess
import sys

METRICS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sectors.sum", "l2_sectors"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("smsp__cycles_active.avg", "smsp_cycles_active"),
]


def to_bytes(val: str, unit: str) -> str:
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    if mult is None or val == "":
        return val
    return str(int(float(val.replace(",", "")) * mult))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), n) for m, n in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch"] + [n + ("_bytes" if n in ("dram_read", "dram_write") else "") for _, n in cols])
        for i, d in enumerate(data):
            vals = []
            for ci, n in cols:
                v = d[ci]
                if n in ("dram_read", "dram_write"):
                    v = to_bytes(v, units[ci])
                if n == "kernel":
                    v = v.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
                vals.append(v)
            w.writerow([i] + vals)
    print(f"wrote {out}: {len(data)} launches")


if __name__ == "__main__":
    main()
