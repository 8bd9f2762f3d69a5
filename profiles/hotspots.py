"""Top stall locations of one kernel from an ncu report:  python profiles/hotspots.py report.ncu-rep [kernel-regex] [n]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "."; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', out)
for blk in blocks[1:]:
    lines = blk.splitlines()
    name = lines[0]
    if not re.search(pat, name):
        continue
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[1:]:
        try:
            n = int(float(r[ci["# Samples"]]))
        except Exception:
            continue
        reasons = sorted(((int(float(r[ci[s]] or 0)), s[6:]) for s in stalls), reverse=True)[:2]
        data.append((n, r[ci["Source"]].strip(), reasons))
    tot = sum(d[0] for d in data) or 1
    print("== %s  (%d samples)" % (name[:90], tot))
    for n, src, reasons in sorted(data, reverse=True)[:top]:
        print("%6d %5.1f%%  %-70s %s" % (n, 100.0 * n / tot, src[:70], " ".join("%s:%d" % (b, a) for a, b in reasons if a)))
    break
