"""Aggregate an `ncu --page source --csv` (SASS view) export by CUDA source line.

usage: python tools/ncu_stalls_by_line.py <sass.csv> <nvdisasm -gi -c listing> <mangled kernel name> [top]
The SASS csv has no line numbers, so instruction offsets are joined against the `//## File ..., line N` annotations of
`nvdisasm -gi -c <cubin>` (cubin from `cuobjdump -xelf all libfse_b200.so`).  Inlined frames are attributed to the
innermost line.
"""
import csv, re, sys, collections

def main():
    sass_csv, listing, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # offset -> (file, line) from the listing
    off2line, cur, inside = {}, None, False
    for ln in open(listing, errors="replace"):
        if ln.startswith("//---") and ".text." in ln:
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass_csv)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr]
    col = {n: i for i, n in enumerate(H)}
    stall_cols = [n for n in H if n.startswith("stall_") and "Not Issued" not in n]
    base = int(rows[hdr + 1][0], 16)
    agg = collections.defaultdict(lambda: collections.Counter())
    total = 0
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        off = int(r[0], 16) - base
        key = off2line.get(off) or ("?", 0)
        n = int(r[col["# Samples"]] or 0)
        total += n
        agg[key]["samples"] += n
        for s in stall_cols:
            v = int(r[col[s]] or 0)
            if v:
                agg[key][s] += v
    print(f"total samples {total}; first instruction offset {min(off2line) if off2line else None:#x} (csv base {base:#x})")
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{k[6:]} {v}" for k, v in c.most_common(5) if k != "samples")
        print(f"{100.0 * c['samples'] / max(total, 1):5.1f}%  {key[0]}:{key[1]:<5d} {st}")

if __name__ == "__main__":
    main()
