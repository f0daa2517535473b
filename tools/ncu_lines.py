"""Aggregate an ncu report's warp-stall samples per CUDA source line (needs -lineinfo + --import-source).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
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
tot = sum(f(r, "# Samples") for r in lines)
print(f"total samples {tot:.0f}")
for r in sorted(lines, key=lambda r: -f(r, "# Samples"))[:topn]:
    print("%5.2f%%  L%-4s long=%-6d short=%-5d wait=%-5d noinst=%-5d math=%-4d exec=%-9d %s" % (
        100 * f(r, "# Samples") / tot, r[0], f(r, "stall_long_sb"), f(r, "stall_short_sb"), f(r, "stall_wait"),
        f(r, "stall_no_inst"), f(r, "stall_math"), f(r, "Instructions Executed"), r[1].strip()[:90]))
