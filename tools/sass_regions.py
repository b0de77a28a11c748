"""Groups the per-instruction counters of an `ncu --page source --csv` export by address range.
usage: python tools/sass_regions.py file.csv [start:end:label ...]   (hex offsets from the kernel start)
Without ranges: prints every instruction with its share of executed warp instructions, average lanes, stall samples."""
import csv, sys

def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    data = []
    base = None
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break                       # the export holds several launches: the first one is taken
        if len(r) < len(hdr) - 5 or r[0] == "Address":
            continue
        a = int(r[ix["Address"]], 16)
        if base is None:
            base = a
        data.append((a - base, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]),
                     int(r[ix["# Samples"]]), int(r[ix["stall_long_sb"]])))
    tot_i = sum(d[2] for d in data); tot_s = sum(d[4] for d in data)
    ranges = []
    for a in [x for x in sys.argv[2:] if not x.startswith("--")]:
        s, e, label = a.split(":")
        ranges.append((int(s, 16), int(e, 16), label))
    if "--blocks" in sys.argv:          # shares per 256-byte block of SASS: enough to tell the node loop, the triangle loop, refill ... apart
        import collections
        blk = collections.OrderedDict()
        for off, src, ie, te, sm, lsb in data:
            b = blk.setdefault(off >> 8, [0, 0, 0, 0, src])
            b[0] += ie; b[1] += te; b[2] += sm; b[3] += lsb
        print("total warp instructions %d, samples %d" % (tot_i, tot_s))
        for k, b in blk.items():
            if b[0]:
                print("%05x  %6.2f%% of warp inst  %5.1f lanes  %6.2f%% of samples (long_sb %5.2f%%)  %s" %
                      (k << 8, 100.0 * b[0] / tot_i, b[1] / max(b[0], 1), 100.0 * b[2] / tot_s, 100.0 * b[3] / tot_s, b[4][:60]))
        return
    if not ranges:
        for off, src, ie, te, sm, lsb in data:
            print("%05x %6.2f%% inst %5.1f lanes %6.2f%% samples  %s" % (off, 100.0 * ie / tot_i, te / max(ie, 1), 100.0 * sm / tot_s, src))
        return
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for s, e, label in ranges:
        sel = [d for d in data if s <= d[0] < e]
        ie = sum(d[2] for d in sel); te = sum(d[3] for d in sel); sm = sum(d[4] for d in sel); ls = sum(d[5] for d in sel)
        print("%-28s %4d SASS  %6.2f%% of warp inst  %5.1f lanes  %6.2f%% of samples (long_sb %5.2f%%)" %
              (label, len(sel), 100.0 * ie / tot_i, te / max(ie, 1), 100.0 * sm / tot_s, 100.0 * ls / tot_s))

if __name__ == "__main__":
    main()
