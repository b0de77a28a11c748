"""Per-launch table from an `ncu --csv --metrics ...` launch list (tools/profile_round2.sh)."""
import csv, collections, sys

def load(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == 'ID':
            hdr, start = r, i
            break
    idx = {n: i for i, n in enumerate(hdr)}
    data = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        key = (int(r[idx['ID']]), r[idx['Kernel Name']])
        data.setdefault(key, {})[r[idx['Metric Name']]] = r[idx['Metric Value']]
    return data

def main():
    data = load(sys.argv[1])
    lim = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9
    for k, v in list(data.items())[:lim]:
        def g(n):
            try:
                return float(v.get(n, '').replace(',', ''))
            except ValueError:
                return float('nan')
        print(k[0], k[1][:28].ljust(28), "t=%7.0fus" % (g('gpu__time_duration.sum') / 1000),
              "lanes=%4.1f" % g('smsp__thread_inst_executed_per_inst_executed.ratio'),
              "issue=%2.0f" % g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
              "l1pipe=%2.0f" % g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
              "lts=%2.0f" % g('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
              "l1hit=%2.0f" % g('l1tex__t_sector_hit_rate.pct'), "l2hit=%2.0f" % g('lts__t_sector_hit_rate.pct'),
              "dram=%5.2fGB" % ((g('dram__bytes_read.sum') + g('dram__bytes_write.sum')) / 1e9),
              "winst=%5.2fG" % (g('smsp__inst_executed.sum') / 1e9),
              "occ=%2.0f" % g('sm__warps_active.avg.pct_of_peak_sustained_active'))

if __name__ == '__main__':
    main()
