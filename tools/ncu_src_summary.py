"""Condense an `ncu --page source --csv` dump (SASS view): per-warp executed instructions and stall samples by opcode and by
blocks of consecutive instructions.   python tools/ncu_src_summary.py gpurun_out/x_src.csv [block]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blk_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ie, isrc, iss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seq = []
for r in rows[2:]:
    if len(r) <= ie:
        continue
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    st = {hdr[i]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
    seq.append((int(r[ie]), int(r[iss]), op, r[isrc].strip(), st))
W = seq[0][0]
tot_s = sum(s[1] for s in seq)
print(f"warps {W}; executed per warp {sum(s[0] for s in seq) / W:.1f}; static {len(seq)}; samples {tot_s}")
ops, samp = collections.Counter(), collections.Counter()
for n, s, op, _, _ in seq:
    ops[op] += n
    samp[op] += s
for op, n in ops.most_common(18):
    print(f"  {op:10s} {n / W:7.1f}/warp  samples {100 * samp[op] / tot_s:5.1f}%")
for i in range(0, len(seq), blk_n):
    b = seq[i:i + blk_n]
    ex, sm = sum(x[0] for x in b) / W, sum(x[1] for x in b)
    if sm == 0 and ex == 0:
        continue
    st = collections.Counter()
    for x in b:
        st.update(x[4])
    top = ", ".join(f"{k[6:]} {v}" for k, v in st.most_common(3))
    hot = max(b, key=lambda x: x[1])
    print(f"{i:5d} exec/warp {ex:6.1f} samples {100 * sm / tot_s:5.1f}%  [{top}]  hot: {hot[3][:60]} ({hot[1]})")
