"""Prints the last N launches (name, grid, duration) of an `ncu --metrics gpu__time_duration.sum --csv --log-file F` capture."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
for r in rows[-int(sys.argv[2]) if len(sys.argv) > 2 else 0:]:
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(r["Metric Unit"], 1)
    print("%9.1f us  %-44s grid %s" % (v, r["Kernel Name"][:44], r["Grid Size"]))
