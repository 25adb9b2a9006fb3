"""Per-kernel DRAM traffic per launch out of an ncu summary CSV (tools/ncu_summary.py output):
  python tools/ncu_traffic.py profiles/<summary>.csv [more.csv ...] > profiles/ncu_traffic.json
bench.py copies `dram_bytes_per_launch` of the dominant kernel into roofline.traffic."""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main(paths):
    acc = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        rd, wr, tm = ix["dram__bytes_read.sum"], ix["dram__bytes_write.sum"], ix["gpu__time_duration.sum"]
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]
            a = acc.setdefault(name, [0, 0.0, 0.0, path])
            a[0] += 1
            a[1] += float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
            a[2] += float(r[tm]) * UNIT[units[tm]]
    # the roofline's kernel CLASS: every tcgen05 3x3 forward / data-gradient kernel of the step (halo, dw-merged, row-strip)
    cls = [v for k, v in acc.items() if k.startswith("tc_conv3")]
    if cls:
        acc["tc_conv3"] = [sum(v[0] for v in cls), sum(v[1] for v in cls), sum(v[2] for v in cls), cls[0][3]]
    out = {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0], "avg_us_under_ncu": v[2] / v[0], "source": v[3]}
           for k, v in sorted(acc.items())}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1:])
