"""Turn the .ncu-rep files / launch list of one round into the small CSV / markdown summaries committed under profiles/."""
import collections, csv, glob, gzip, os, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEEP = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg.per_second', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active']
out_md = [f"# ncu summaries, {R}\n"]
for rep in sorted(glob.glob(f"gpurun_out/{R}_*.ncu-rep")):
    name = os.path.basename(rep)[len(R) + 1:-8]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEEP]
    with open(f"profiles/{R}_ncu_{name}_summary.csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    out_md.append(f"## {name}\n")
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        g = lambda k: r[col[k]] if k in col else "n/a"
        u = lambda k: units[col[k]] if k in col else ""
        out_md.append(f"- `{g('Kernel Name')[:70]}` grid {g('Grid Size')}: {g('gpu__time_duration.sum')} {u('gpu__time_duration.sum')}, "
                      f"DRAM r/w {g('dram__bytes_read.sum')} {u('dram__bytes_read.sum')} / {g('dram__bytes_write.sum')} {u('dram__bytes_write.sum')}, "
                      f"DRAM {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} %, tensor pipe {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} %, "
                      f"regs {g('launch__registers_per_thread')}, warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active')} %")
    out_md.append("")
ll = f"gpurun_out/{R}_launches.csv"
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) <= mi:
            continue
        nm = r[ki].split('(')[0].replace('ovla::', '').replace('void ', '')
        v = float(r[mi].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(r[ui], v)
        agg[nm][0] += 1; agg[nm][1] += v
    tot = sum(v[1] for v in agg.values())
    out_md.append(f"## launch list of one bs=256 step ({sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms cold-cache serialised)\n")
    out_md.append("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out_md.append(f"| `{k[:80]}` | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / tot:.2f} % |")
    with open(ll, "rb") as fi, gzip.open(f"profiles/{R}_ncu_launches.csv.gz", "wb") as fo:
        shutil.copyfileobj(fi, fo)
open(f"profiles/{R}_ncu_summary.md", "w").write("\n".join(out_md) + "\n")
print("\n".join(out_md))
