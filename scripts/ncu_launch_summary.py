"""Turn an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...``) into the two files kept
under profiles/: a trimmed CSV (id, kernel, grid, block, duration) and a per-kernel summary JSON with shares.

    python scripts/ncu_launch_summary.py gpurun_out/r2_ncu_launches_c3.csv profiles/r2_ncu_launches_c3 "<command>" "<note>"
"""
import collections
import csv
import json
import re
import sys


def main(src: str, dst_prefix: str, command: str, note: str) -> None:
    lines = [ln for ln in open(src, errors="replace") if not ln.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
        rows.append((row["ID"], row["Kernel Name"], row.get("Grid Size", ""), row.get("Block Size", ""), us))
    with open(dst_prefix + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration_us"])
        for r in rows:
            w.writerow([r[0], r[1][:140], r[2], r[3], f"{r[4]:.3f}"])
    tot = collections.defaultdict(lambda: [0, 0.0])
    for _, name, _, _, us in rows:
        n = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
        n = re.sub(r"\(.*$", "", n)
        tot[n][0] += 1
        tot[n][1] += us
    total = sum(v[1] for v in tot.values())
    gemm = sum(v[1] for k, v in tot.items() if "zgemm" in k)
    out = {"command": command, "note": note, "launches": len(rows), "total_us": round(total, 1),
           "zgemm_share": round(gemm / total, 4) if total else None,
           "kernels": {k: {"launches": c, "total_us": round(t, 1), "share": round(t / total, 4), "avg_us": round(t / c, 2)}
                       for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])}}
    json.dump(out, open(dst_prefix + "_summary.json", "w"), indent=1)
    print(len(rows), "launches", round(total / 1000, 2), "ms; zgemm share", out["zgemm_share"])
    for k, v in list(out["kernels"].items())[:12]:
        print(f"  {v['share']:.3f} {v['launches']:6d} {v['avg_us']:10.2f} us  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "", sys.argv[4] if len(sys.argv) > 4 else "")
