"""Summarise .ncu-rep captures into a small markdown table (run in the build container; needs `ncu`).

    python tools/summarize_ncu.py gpurun_out/prof_*.ncu-rep > profiles/r01_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def stalls(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    res = []
    start = 0
    while start < len(rows):
        if rows[start] and rows[start][0] == "Kernel Name":
            hdr = rows[start + 1]
            data = []
            i = start + 2
            while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
                data.append(rows[i])
                i += 1
            cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            agg = {c: 0 for c in cols}
            for r in data:
                for c in cols:
                    try:
                        agg[c] += int(r[hdr.index(c)])
                    except (ValueError, IndexError):
                        pass
            tot = sum(agg.values()) or 1
            res.append(", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:4]))
            start = i
        else:
            start += 1
    return res


def main():
    print("| capture | kernel | " + " | ".join(k for _, k in KEYS) + " | top warp stalls |")
    print("|---|---|" + "---|" * (len(KEYS) + 1))
    for path in sys.argv[1:]:
        recs, units = raw(path)
        st = stalls(path)
        for n, r in enumerate(recs):
            cells = []
            for key, _ in KEYS:
                v = r.get(key, "")
                u = units.get(key, "")
                cells.append(f"{v} {u}".strip())
            name = r.get("Kernel Name", "?").split("(")[0][-40:]
            print(f"| {path.split('/')[-1]} | {name} | " + " | ".join(cells) + f" | {st[n] if n < len(st) else ''} |")


if __name__ == "__main__":
    main()
