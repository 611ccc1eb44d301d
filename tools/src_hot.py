"""Hot source lines of one kernel: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:X > f.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
cur, out, idx = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
    elif r[0] == "Line No":
        idx = {k: i for i, k in enumerate(r)}
    elif r[0].isdigit() and idx and cur:
        try:
            out.append((int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]]), cur.split("/")[-1], int(r[0]), r[1].strip()[:100]))
        except Exception:
            pass
S = sum(o[0] for o in out); I = sum(o[1] for o in out)
print("samples", S, "warp-instr", I)
for s, i, f, l, src in sorted(out, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%5.1f%% smp %5.1f%% ins  %s:%-4d %s" % (100 * s / max(S, 1), 100 * i / max(I, 1), f, l, src))
