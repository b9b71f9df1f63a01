"""Per-source-line instruction / stall-sample shares from an .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if len(r) > 10 and r[0] not in ("", "Line No"):
        try: agg.append((cur_file, int(r[0]), r[1], int(r[6]), int(r[7]), int(r[8])))
        except ValueError: pass
tot_s = sum(a[3] for a in agg) or 1; tot_i = sum(a[4] for a in agg) or 1
print("total samples", tot_s, "total warp inst", tot_i)
agg.sort(key=lambda a: -(a[4] / tot_i + a[3] / tot_s))
for a in agg[:top]:
    print("%-14s %4d samp %5.1f%% inst %5.1f%% thr/inst %4.1f | %s" % (a[0][:14], a[1], 100 * a[3] / tot_s, 100 * a[4] / tot_i, a[5] / max(1, a[4]), a[2].strip()[:100]))
