"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / cluster use in the shipped library:
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
(cuobjdump -sass on openvla_probe_b200/libovla_b200.so; runs without a GPU)."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "openvla_probe_b200", "libovla_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "UBLKCP", "UBLKPF",
       "SYNCS", "HMMA", "FFMA2", "MUFU", "ACQBULK", "CCTL", "ERRBAR", "UCGABAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
per, cur, tot = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    per[cur]["_all"] += 1
    for p in PAT:
        if op.startswith(p):
            key = p
            if p in ("UTCHMMA", "UTMALDG", "UTMASTG"):
                key = ".".join(op.split(".")[:2]) if "." in op else p
                if ".2CTA" in op:
                    key += "(2CTA)"
            per[cur][key] += 1
            tot[key] += 1
            break
print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a cubin): instruction counts per kernel")
print("# totals: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items())))
for k, c in per.items():
    feats = ", ".join(f"{n} {v}" for n, v in sorted(c.items()) if n != "_all")
    print(f"{k} :: {c['_all']} instr :: {feats}")
