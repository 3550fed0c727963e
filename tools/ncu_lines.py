"""Executed instructions and stall samples per CUDA source line of one kernel: joins `ncu --page source --csv` (per SASS
instruction) with `nvdisasm -g` line markers of the object the kernel was built from (both list the instructions of the
kernel and of its out-of-line callees in the same order).

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled-name substring> [top N] [warps for the per-warp column] [launch index]
"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, obj, name = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
nw = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
sec = int(sys.argv[6]) if len(sys.argv) > 6 else 0
rows = rows[starts[sec]:starts[sec + 1]]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [(int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), r[ix["Source"]].strip())
        for r in rows[2:] if len(r) >= len(hdr)]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and name in l)
ins, cur, sec = [], None, None
for l in dis[start:]:
    if l.startswith("//--------------------- .text."):
        nm = l.split(".text.")[1].split(" ")[0]
        if sec is not None and "kernel" in nm and name not in nm and len(ins) >= len(data):
            break
        sec = nm
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(1).strip()))
ins = ins[:len(data)]
assert len(ins) == len(data), (len(ins), len(data))
opc = lambda t: [w for w in t.replace("{", " ").split() if not w.startswith("@")][0].split(".")[0]
bad = sum(1 for (_, a), (_, _, b) in zip(ins, data) if opc(a) != opc(b))
if bad:
    print("WARNING: %d of %d instructions do not match the object (another build?): the attribution below is unreliable" % (bad, len(data)))
ins = [c for c, _ in ins]
by, bs = collections.Counter(), collections.Counter()
for line, (n, s, _) in zip(ins, data):
    by[line] += n
    bs[line] += s
tot, tots = sum(by.values()), max(sum(bs.values()), 1)
print("%s: %d static instructions, %.0f executed per warp" % (rows[0][1][:70], len(data), tot / nw))
for line, n in by.most_common(top):
    print("%-14s %5d  %8.0f /warp  %5.1f %% of instructions  %5.1f %% of samples" % (line[0], line[1], n / nw, 100.0 * n / tot, 100.0 * bs[line] / tots))
