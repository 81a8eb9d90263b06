"""profiles/traffic.json from an ncu metrics pass of bench.py (roofline.traffic of the bench line).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-config2 --no-profile
    python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/traffic.json

Per kernel class: DRAM bytes (read + write) per launch, averaged over every launch in the capture, and the launch count.
"""
import csv
import json
import sys

CLASSES = {
    "conv": ("conv_halo_kernel", "conv_igemm_kernel", "conv_prog_kernel"),
    "corr_lookup": ("corr_lookup",),
    "imgprop": ("imgprop_persistent",),
    "dcn_sample": ("dcn_sample",),
    "featprop_warp": ("featprop_cond",),
    "fold_ffn": ("fold7x7s3",),
    "attention": ("window_attention",),
}


def main(src, dst):
    rows = []
    with open(src, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per_id = {}
    for r in rd:
        k = per_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
        try:
            k[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
            k[r["Metric Name"] + ".unit"] = r["Metric Unit"]
        except ValueError:
            pass
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {"note": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch from " + src +
                   " (ncu replays each kernel with cold caches; average over all launches of the class)"}
    for cls, pats in CLASSES.items():
        tot, n, ns = 0.0, 0, 0.0
        for k in per_id.values():
            if any(p in k["name"] for p in pats):
                b = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    if m in k:
                        b += k[m] * scale.get(k.get(m + ".unit", "byte"), 1.0)
                tot += b
                n += 1
                ns += k.get("gpu__time_duration.sum", 0.0)
        if n:
            out[cls + "_bytes_per_launch"] = tot / n
            out[cls + "_launches"] = n
            out[cls + "_total_gbytes"] = tot / 1e9
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
