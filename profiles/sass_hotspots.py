#!/usr/bin/env python
"""Join an ncu source-page CSV (SASS level, `ncu -i X.ncu-rep --page source --csv`) with `nvdisasm -gi` line info of
the same kernel, and print the warp-stall samples per source line (innermost line + inlining call sites).

    python profiles/sass_hotspots.py src.csv kernel.sass [top]
"""
import collections
import csv
import re
import sys


def line_map(sass_path):
    """offset -> ['file:line' innermost, ..., outermost call site] from nvdisasm -gi output (one comment line per
    inlining level, innermost first, in front of the instructions it covers)."""
    cur, fresh = [], True
    out = {}
    for ln in open(sass_path):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                cur, fresh = [], False
            cur.append(f"{m.group(1).split('/')[-1].replace('gemm_sm100.cuh', 'gemm')}:{m.group(2)}")
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
            fresh = True
    return out


def main():
    src, sass = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    lm = line_map(sass)
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    ia, isamp = hdr.index('Address'), hdr.index('# Samples')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    base = min(int(r[ia], 16) for r in data)
    agg = collections.defaultdict(lambda: [0, collections.Counter()])
    tot = 0
    for r in data:
        n = int(r[isamp] or 0)
        if not n:
            continue
        tot += n
        off = int(r[ia], 16) - base
        chain, _ = lm.get(off, (["?"], ""))
        key = " <- ".join(chain)
        agg[key][0] += n
        for i in stall:
            v = int(r[i] or 0)
            if v:
                agg[key][1][hdr[i]] += v
    print(f"# {tot} samples")
    for k, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{n:6d} {100.0 * n / tot:5.1f}%  {k}   {dict(st.most_common(3))}")


if __name__ == "__main__":
    main()
