import csv, re, sys
src_csv, sass, kname = sys.argv[1:4]
rows = list(csv.reader(open(src_csv))); hdr = rows[1]
iexec = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); isamp = hdr.index("# Samples")
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = None; seq = []
loc_re = re.compile(r'//## File "([^"]+)", line (\d+)'); ins_re = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);')
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."): break
    m = loc_re.search(l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = ins_re.match(l)
    if m: seq.append(cur)
for i, r in enumerate(rows[2:]):
    if len(r) > iexec:
        print(f"{i:5d} {int(r[iexec]):11d} {int(r[isamp]):6d} {seq[i][0][:10] if seq[i] else '-':10s}:{seq[i][1] if seq[i] else 0:<4d} {r[isrc].strip()}")
