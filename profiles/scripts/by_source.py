"""Per-source-line view of an `ncu --set full --import-source on` capture: executed warp instructions and stall samples per source line
(top lines) and per region of the source file, joined from ncu's per-SASS-instruction page and `nvdisasm --print-line-info` of the same build.
usage: by_source.py <ncu --page source --csv> <nvdisasm.sass> <mangled kernel> <source file> "title" [line:label ...]"""
import collections, csv, re, sys
src_csv, sass, kname, srcfile, title = sys.argv[1:6]
marks = sorted((int(a.split(":")[0]), a.split(":", 1)[1]) for a in sys.argv[6:])
rows = list(csv.reader(open(src_csv))); hdr = rows[1]
iexec = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
insts = [(int(r[iexec]), int(r[isamp])) for r in rows[2:] if len(r) > iexec]
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
loc_re = re.compile(r'//## File "([^"]+)", line (\d+)'); ins_re = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);')
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."): break
    m = loc_re.search(l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if ins_re.match(l): seq.append(cur)
n = min(len(insts), len(seq))
T = sum(i[0] for i in insts[:n]); S = sum(i[1] for i in insts[:n])
ex = collections.Counter(); sm = collections.Counter(); st = collections.Counter()
for i in range(n):
    ex[seq[i]] += insts[i][0]; sm[seq[i]] += insts[i][1]; st[seq[i]] += 1
text = open(srcfile).read().split("\n"); base = srcfile.split("/")[-1]
print(title)
print(f"SASS instructions {n} (executed at least once: {sum(1 for i in insts[:n] if i[0] > 0)}), warp instructions executed {T:.4g}, stall samples {S}")
if marks:
    print("-- by region of", base, "(inlined code of other files is attributed to where it was inlined FROM: see the top lines)")
    for j, (ln, label) in enumerate(marks):
        hi = marks[j + 1][0] if j + 1 < len(marks) else 10 ** 9
        e = sum(v for k, v in ex.items() if k and k[0] == base and ln <= k[1] < hi); s_ = sum(v for k, v in sm.items() if k and k[0] == base and ln <= k[1] < hi)
        print(f"  lines {ln:5d}-{(hi - 1 if hi < 10**9 else len(text)):5d}  instr {100 * e / T:5.1f} %  samples {100 * s_ / max(S, 1):5.1f} %  {label}")
    e = sum(v for k, v in ex.items() if not k or k[0] != base); s_ = sum(v for k, v in sm.items() if not k or k[0] != base)
    print(f"  other files (common.cuh: Philox, warp reductions; CUDA intrinsics)  instr {100 * e / T:5.1f} %  samples {100 * s_ / max(S, 1):5.1f} %")
print("-- top source lines")
for k, v in ex.most_common(30):
    t = text[k[1] - 1].strip()[:100] if k and k[0] == base and k[1] - 1 < len(text) else ""
    print(f"  {100 * v / T:5.2f} %  samples {100 * sm[k] / max(S, 1):5.2f} %  static {st[k]:4d}  {k[0] if k else '-'}:{k[1] if k else 0}  {t}")
