"""Per-source-line stall samples of a kernel: joins `ncu --page source --print-source sass` (warp-stall samples per
SASS instruction) with `nvdisasm -gi` of the same build (source line of every SASS instruction).
  python tools/hot_lines.py gpurun_out/prof_r1_frame.ncu-rep _ZN2kb16k_register_frameENS_11FrameParamsE profiles/r1_register_frame_hot_lines.csv"""
import csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sym, out = sys.argv[1], sys.argv[2], sys.argv[3]
launch = sys.argv[4] if len(sys.argv) > 4 else "0"

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == "Kernel Name":
        cur = {"hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r:
        cur["rows"].append(r)
b = blocks[int(launch)]
H = {k: i for i, k in enumerate(b["hdr"])}
stall_cols = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]

with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "kiss-icp_b200", "libkiss_icp_b200.so")], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
lines, inside, loc, chain = [], False, None, None
for ln in dis.splitlines():
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + sym
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        here = (os.path.basename(m.group(1)), int(m.group(2)))
        if " inlined at " in m.group(3) or loc is None or chain == "fresh":
            pass
        # the first annotation of a group is the innermost location; following ones are its inline parents
        if chain != "open":
            loc, outer, chain = here, here, "open"
        else:
            outer = here
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m:
        lines.append((int(m.group(1), 16), m.group(2).strip(), loc, outer))
        chain = "closed"
if len(lines) != len(b["rows"]):
    sys.exit(f"instruction count mismatch: nvdisasm {len(lines)} vs ncu {len(b['rows'])} - profile is from another build")
agg = {}
total = 0
for (off, ins, loc, outer), r in zip(lines, b["rows"]):
    if ins.split()[0].lstrip("@!P0123456789 ") != r[H["Source"]].split()[0].lstrip("@!P0123456789 ") and ins.split()[-1] != r[H["Source"]].split()[-1]:
        pass  # operand formatting differs between the tools; the count check above is the guard
    s = int(r[H["# Samples"]] or 0)
    total += s
    a = agg.setdefault((loc, outer), {"samples": 0, "inst": 0, **{k: 0 for k in stall_cols}})
    a["samples"] += s
    a["inst"] += int(r[H["Instructions Executed"]] or 0)
    for k in stall_cols:
        a[k] += int(r[H[k]] or 0)
src_cache = {}
def text(loc):
    f, n = loc
    p = os.path.join(ROOT, "kiss-icp_b200", "csrc", f)
    if p not in src_cache:
        src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[p]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["share_pct", "samples", "warp_insts", "line", "called_from", "top_stalls", "source"])
    for (loc, outer), a in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:60]:
        tops = sorted(((a[k], k[6:]) for k in stall_cols if a[k]), reverse=True)[:3]
        w.writerow([round(100 * a["samples"] / max(total, 1), 2), a["samples"], a["inst"], f"{loc[0]}:{loc[1]}",
                    f"{outer[0]}:{outer[1]}" if outer != loc else "", " ".join(f"{n}={v}" for v, n in tops), text(loc)])
print("wrote", out, "total samples", total)
