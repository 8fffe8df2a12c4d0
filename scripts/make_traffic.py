"""profiles/r2_traffic.json from an `ncu --set full` capture of the dominant kernel, taken on THIS build.

  python scripts/make_traffic.py gpurun_out/r2_prof_wgs30x.ncu-rep [workload]

Reads the report with `ncu -i ... --page raw --csv`, takes dram__bytes_read.sum + dram__bytes_write.sum of every
k_depth_tiles16 launch in it (per launch, like roofline.achieved) and stores the mean together with the hash of the
kernel's sources (bench.kernel_source_sha): bench.py only quotes the file when the hash still matches the sources it was
built from, and prints traffic: null otherwise.  Also writes the raw page beside it (profiles/r2_ncu_full_<workload>_raw.csv)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1)


def main():
    rep = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "wgs30x"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    raw_path = os.path.join(ROOT, "profiles", "r2_ncu_full_%s_raw.csv" % workload)
    open(raw_path, "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(units[col["gpu__time_duration.sum"]], 1e-6)
        out.setdefault(name, []).append((rd, wr, dur_ms))
    summary = {k: {"launches": len(v), "dram_read_bytes": sum(x[0] for x in v) / len(v), "dram_write_bytes": sum(x[1] for x in v) / len(v),
                   "duration_ms": sum(x[2] for x in v) / len(v)} for k, v in out.items()}
    tiles = next((v for k, v in summary.items() if "k_depth_tiles16" in k), None)
    if tiles is None:
        raise SystemExit("no k_depth_tiles16 launch in " + rep)
    doc = {"workload": workload, "kernel": "k_depth_tiles16", "kernel_source_sha": bench.kernel_source_sha(),
           "traffic_bytes_per_launch": tiles["dram_read_bytes"] + tiles["dram_write_bytes"],
           "dram_read_bytes": tiles["dram_read_bytes"], "dram_write_bytes": tiles["dram_write_bytes"], "duration_ms_under_ncu": tiles["duration_ms"],
           "capture": "ncu --set full --clock-control none, %s, %d launch(es)" % (os.path.basename(rep), tiles["launches"]), "kernels": summary}
    json.dump(doc, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    print(json.dumps({k: doc[k] for k in ("kernel_source_sha", "traffic_bytes_per_launch", "dram_read_bytes", "dram_write_bytes", "duration_ms_under_ncu")}))


if __name__ == "__main__":
    main()
