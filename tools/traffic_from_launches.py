"""Aggregate an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch of one
denoiser call) into profiles/traffic.json: average DRAM bytes and time per launch for every kernel class.
    python tools/traffic_from_launches.py gpurun_out/launches_r1.csv profiles/traffic.json "<command that produced it>"
"""
import collections
import csv
import json
import sys

CLASSES = [("gemm_tc", "gemm_tc"), ("chain", "chain_tc"), ("attention", "attention"), ("gn_apply", "gn_apply"),
           ("layernorm", "layernorm"), ("axpbypcz", "sampler_update"), ("temb", "time_embedding"),
           ("nct_to_ntc", "layout")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "nsecond": 1e-3, "msecond": 1e3}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    how = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[1:]:
        val = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        per[r[ii]][r[mi]] = val
        names[r[ii]] = r[ki]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for lid, m in per.items():
        cls = next((c for k, c in CLASSES if k in names[lid]), "other")
        a = agg[cls]
        a[0] += 1
        a[1] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a[2] += m.get("gpu__time_duration.sum", 0.0)
    total_t = sum(a[2] for a in agg.values())
    out = {"source": how, "launches": {k: a[0] for k, a in agg.items()},
           "per_launch_dram_bytes": {k: round(a[1] / a[0]) for k, a in agg.items()},
           "per_launch_us": {k: round(a[2] / a[0], 3) for k, a in agg.items()},
           "share_of_call_time": {k: round(a[2] / total_t, 4) for k, a in agg.items()}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
