#!/usr/bin/env python
"""Builds profiles/ncu_traffic.json and the per-capture summaries profiles/r2_ncu_<tag>.txt from the raw CSV exports of
profiles/capture_r2.sh (gpurun_out/r2_ncu_<tag>.raw.csv).  traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch."""
import csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
POINTS = {  # tag -> (workload, step index of the captured launches, capture command)
    "drumA": ("drum", 15, "ncu --set full --clock-control none -k regex:k_contact_sphere_pp|k_gather -s 30 -c 2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-poly"),
    "drumB": ("drum", 220, "ncu --set full --clock-control none -k regex:k_contact_sphere_pp|k_gather -s 440 -c 2 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-e2e --no-poly"),
    "polystep": ("poly", 15, "ncu --set full --clock-control none -k regex:k_detect_poly_pp|k_law_poly_pp|k_gather -s 45 -c 3 python bench.py --workload poly --particles 500000 --steps 20 --warmup 5 ..."),
    "polynbh": ("poly", 0, "ncu --set full --clock-control none -k regex:k_nbh_poly_cell|k_nbh_poly_replay -s 1 -c 2 python bench.py --workload poly --particles 500000 --steps 6 --warmup 3 ..."),
    "sphnbh": ("drum", 0, "ncu --set full --clock-control none -k regex:k_nbh_sphere -s 2 -c 2 python bench.py --steps 6 --warmup 3 ..."),
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def to_bytes(v, unit):
    m = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(unit, 1.0)


def main():
    points = []
    for tag, (wl, step, cmd) in POINTS.items():
        raw = OUT / f"r2_ncu_{tag}.raw.csv"
        if not raw.exists():
            continue
        rows = list(csv.reader(open(raw)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        kernels = {}
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            short = name.split("(")[0].replace("void ", "").strip()
            key = short.split("<")[0]
            rd = to_bytes(num(r[ix["dram__bytes_read.sum"]]), units[ix["dram__bytes_read.sum"]])
            wr = to_bytes(num(r[ix["dram__bytes_write.sum"]]), units[ix["dram__bytes_write.sum"]])
            t = num(r[ix["gpu__time_duration.sum"]])
            tu = units[ix["gpu__time_duration.sum"]]
            t_us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(tu, 1.0)
            kernels.setdefault(key, {"name": short, "traffic": rd + wr, "read": rd, "write": wr, "time_us_under_ncu": t_us})
        if wl == "poly" and "k_detect_poly_pp" in kernels and "k_law_poly_pp" in kernels:
            kernels["k_detect_poly_pp+k_law_poly_pp"] = kernels["k_detect_poly_pp"]["traffic"] + kernels["k_law_poly_pp"]["traffic"]
        flat = {k: (v["traffic"] if isinstance(v, dict) else v) for k, v in kernels.items()}
        points.append({"tag": tag, "workload": wl, "step_index": step, "capture": cmd, "kernels": flat,
                       "detail": {k: v for k, v in kernels.items() if isinstance(v, dict)}})
        txt = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(raw)], capture_output=True, text=True).stdout
        (ROOT / "profiles" / f"r2_ncu_{tag}.txt").write_text(f"# {cmd}\n# workload {wl}, launches of step ~{step}\n" + txt)
    (ROOT / "profiles" / "ncu_traffic.json").write_text(json.dumps({"points": points, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"}, indent=1))
    for p in points:
        print(p["tag"], {k: round(v / 1e6, 1) for k, v in p["kernels"].items()}, "MB")


if __name__ == "__main__":
    main()
