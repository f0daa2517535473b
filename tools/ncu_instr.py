"""Dynamic warp-instruction counts per CUDA source line from an ncu report (where do the instructions go?)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and r and r[0] not in ("", "File Path", "Function Name"):
        lines.append(r)
ix = {n: i for i, n in enumerate(hdr)}
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
tot = sum(f(r, "Instructions Executed") for r in lines)
print(f"total warp instructions {tot:.0f}")
for r in sorted(lines, key=lambda r: -f(r, "Instructions Executed"))[:topn]:
    print("%5.2f%%  L%-4s %s" % (100 * f(r, "Instructions Executed") / tot, r[0], r[1].strip()[:110]))
