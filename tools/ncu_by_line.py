#!/usr/bin/env python
"""Join an `ncu --page source --print-source sass --csv` dump with `nvdisasm -g`
line info and aggregate stall samples / executed instructions per source line and
per enclosing device function.
usage: ncu_by_line.py <sass.csv> <nvdisasm -g listing> <kernel mangled-name substring> [top]"""
import csv, re, sys, collections, os

def parse_listing(path, kern):
    off2line, cur, active = {}, None, False
    for ln in open(path, errors="ignore"):
        if ln.startswith("\t.section\t.text."):
            active = kern in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return off2line

def func_table(srcdir):
    tab = {}
    for f in os.listdir(srcdir):
        if not f.endswith((".cuh", ".cu")):
            continue
        starts = []
        for i, ln in enumerate(open(os.path.join(srcdir, f)), 1):
            m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|inline|static).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", ln)
            if m and not ln.startswith(" "):
                starts.append((i, m.group(1)))
        tab[f] = starts
    return tab

def func_of(tab, key):
    if key is None:
        return "?"
    f, line = key
    name = "?"
    for s, n in tab.get(f, []):
        if s <= line:
            name = n
    return name

def main():
    sass_csv, listing, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    off2line = parse_listing(listing, kern)
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ia, isamp, iinst, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][ia], 16)
    by_line = collections.defaultdict(lambda: [0.0, 0.0])
    by_func = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    stalls = collections.defaultdict(float)
    fstall = collections.defaultdict(lambda: collections.defaultdict(float))
    tab = func_table(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rsr_mjx_b200", "csrc"))
    tot_s = tot_i = 0.0
    for r in data:
        off = int(r[ia], 16) - base
        key, _ = off2line.get(off, (None, ""))
        s, n = float(r[isamp] or 0), float(r[iinst] or 0)
        tot_s += s; tot_i += n
        by_line[key][0] += s; by_line[key][1] += n
        fn = func_of(tab, key)
        by_func[fn][0] += s; by_func[fn][1] += n; by_func[fn][2] += n * float(r[ithr] or 0)
        for i, h in stall_cols:
            stalls[h] += float(r[i] or 0)
            fstall[h][fn] += float(r[i] or 0)
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.3e}")
    print("--- by function (samples%, inst%, avg active threads)")
    for fn, (s, n, t) in sorted(by_func.items(), key=lambda kv: -kv[1][0]):
        if s / tot_s > 0.002:
            print(f"  {100*s/tot_s:5.1f}%  {100*n/tot_i:5.1f}%  {t/max(n,1):5.1f}  {fn}")
    print("--- stall reasons (all samples)")
    ts = sum(stalls.values())
    for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {100*v/ts:5.1f}%  {h}")
    for h in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving"):
        tot = sum(fstall[h].values()) or 1.0
        best = sorted(fstall[h].items(), key=lambda kv: -kv[1])[:8]
        print(f"--- {h} by function: " + ", ".join(f"{fn} {100*v/tot:.0f}%" for fn, v in best))
    print(f"--- top {top} lines")
    for key, (s, n) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100*s/tot_s:5.2f}% samp {100*n/tot_i:5.2f}% inst  {key}")

if __name__ == "__main__":
    main()
