"""Join ncu's per-SASS-instruction counts with nvdisasm line info -> per-source-line instruction totals.
usage: sass_lines.py <ncu source csv> <all.sass> <mangled kernel name> [top N]"""
import csv, re, sys, collections
src_csv, sass, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iexec = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); isamp = hdr.index("# Samples")
texec = hdr.index("Thread Instructions Executed")
wave = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
insts = [(r[isrc].strip(), int(r[iexec]), int(r[isamp]), int(r[wave]) if wave is not None and r[wave] not in ("", "-") else 0) for r in rows[2:] if len(r) > iexec]
# nvdisasm: find function section, walk lines
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = None; seq = []
loc_re = re.compile(r'//## File "([^"]+)", line (\d+)')
inl_re = re.compile(r'inlined at "([^"]+)", line (\d+)')
ins_re = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);')
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = loc_re.search(l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = ins_re.match(l)
    if m:
        seq.append((cur, m.group(1)))
print("ncu insts", len(insts), "nvdisasm insts", len(seq), file=sys.stderr)
n = min(len(insts), len(seq))
tot = collections.Counter(); samp = collections.Counter(); wv = collections.Counter()
for i in range(n):
    tot[seq[i][0]] += insts[i][1]; samp[seq[i][0]] += insts[i][2]; wv[seq[i][0]] += insts[i][3]
T = sum(tot.values()); S = sum(samp.values())
print(f"total warp-instr {T:.4g}, samples {S}")
for k, v in tot.most_common(topn):
    print(f"{str(k):32s} instr {v:12d} {100*v/T:5.1f}%  samples {100*samp[k]/max(S,1):5.1f}%  smem wavefronts {wv[k]}")
if len(sys.argv) > 5:
    print("---- by range")
    rs = [tuple(map(int, x.split("-"))) for x in sys.argv[5].split(",")]
    for a, b in rs:
        v = sum(c for k, c in tot.items() if k and k[0] == "prefix.cuh" and a <= k[1] <= b)
        w = sum(c for k, c in wv.items() if k and k[0] == "prefix.cuh" and a <= k[1] <= b)
        print(f"prefix.cuh {a}-{b}: {100*v/T:5.1f}%  wavefronts {w}")
    v = sum(c for k, c in tot.items() if not k or k[0] != "prefix.cuh")
    print(f"other files: {100*v/T:5.1f}%")
    for f in set(k[0] for k in tot if k):
        print(f, f"{100*sum(c for k, c in tot.items() if k and k[0]==f)/T:5.1f}%")
