"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line (executed warp instructions + stall samples).
   python tools/ncu_hot_lines.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file = None; hdr = None
agg = collections.Counter(); samples = collections.Counter(); text = {}; opc = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        samp = next((i for i, h in enumerate(r) if h == "# Samples"), None)
        if samp is None: samp = next((i for i, h in enumerate(r) if "Sampling" in h and "All" in h), None)   # column name differs between ncu versions
        continue
    if r[0] == "Function Name" or hdr is None: continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    try:   # source text with quotes / commas (inline asm) can shift the columns of a row: skip what does not parse
        ie = r[hdr["Instructions Executed"]]
        if not ie: continue
        n_ie, n_s = int(float(ie)), (int(float(r[samp] or 0)) if samp is not None else 0)
    except (ValueError, IndexError):
        continue
    key = ((cur_file or "?").split("/")[-1], line)
    agg[key] += n_ie; samples[key] += n_s
    # first "Source" column = CUDA-C text for this row kind, second = SASS
    src_cols = [i for i, h in enumerate(rows[2]) if h == "Source"] if False else None
tot = sum(agg.values())
print("total warp instructions (rows with both views are double counted consistently):", tot)
for (f, l), n in agg.most_common(top):
    print(f"{n:>12} {100.0 * n / tot:5.1f}%  samples {samples[(f, l)]:>6}  {f}:{l}")
