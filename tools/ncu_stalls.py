"""Top stalled SASS instructions of one kernel of an .ncu-rep captured with --import-source on:
    python tools/ncu_stalls.py gpurun_out/r02j_gemm_dino.ncu-rep 1 [n_lines]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 28
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []; blocks.append(cur); continue
    if cur is not None:
        cur.append(line)
rows = list(csv.reader(blocks[kid]))
hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
tot = sum(int(r[col["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:n]:
    st = sorted(((h, int(r[col[h]])) for h in stall_cols if int(r[col[h]]) > 0), key=lambda kv: -kv[1])[:3]
    print(f'{int(r[col["# Samples"]]):6d} {100 * int(r[col["# Samples"]]) / tot:5.1f}%  {r[col["Source"]].strip()[:72]:72s} {st}')
