"""Top SASS lines by warp-stall samples out of an .ncu-rep captured with --import-source on / -lineinfo:
  python tools/ncu_source_top.py rep.ncu-rep [N] > profiles/name_ncu_source_top.txt"""
import csv
import subprocess
import sys


def main(rep, top=30):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kernel, hdr, data = rows[0][1], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    s_all = ix["Warp Stall Sampling (All Samples)"]
    data = [r for r in data if len(r) > s_all and r[s_all].isdigit()]
    total = sum(int(r[s_all]) for r in data)
    print("kernel:", kernel)
    print("warp-stall samples (all):", total, " SASS instructions:", len(data))
    print("%8s %6s  %s" % ("samples", "share", "SASS"))
    for r in sorted(data, key=lambda r: -int(r[s_all]))[:top]:
        print("%8d %5.1f%%  %s" % (int(r[s_all]), 100.0 * int(r[s_all]) / max(total, 1), r[ix["Source"]].strip()))
    # by mnemonic
    by = {}
    for r in data:
        src = r[ix["Source"]].strip()
        parts = src.split()
        m = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
        m = m.split(".")[0]
        by[m] = by.get(m, 0) + int(r[s_all])
    print("\nby mnemonic:")
    for m, v in sorted(by.items(), key=lambda kv: -kv[1])[:15]:
        print("%8d %5.1f%%  %s" % (v, 100.0 * v / max(total, 1), m))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
