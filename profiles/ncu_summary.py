"""Print the headline metrics (and optionally the SASS opcode mix) of an .ncu-rep capture."""
import collections
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_bytes.sum', 'sm__cycles_elapsed.avg']


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"{h} = {vals[i]} {units[i]}")
        print()


def sass(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    ops, samp, st = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    for r in data:
        try:
            n, s = int(r[ie]), int(r[isamp])
        except Exception:
            continue
        t = r[ia].split()
        op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).split('.')[0]
        ops[op] += n
        samp[op] += s
        tot += n
        for i in stall:
            try:
                st[hdr[i]] += int(r[i])
            except Exception:
                pass
    print(f"total warp instructions {tot}")
    for op, n in ops.most_common(18):
        print(f"  {op:8s} {n / tot * 100:5.1f}%  samples {samp[op]}")
    ts = sum(st.values())
    print("stall samples:", ", ".join(f"{k[6:]} {v / ts * 100:.0f}%" for k, v in st.most_common(8)))


def traffic(path, out_json):
    """Write {kernels: {presmooth|postsmooth|apply_p|update_xr: {dram_bytes, time_us, grid}}} for the level-0
    launches (largest grid of each kernel) of a capture; bench.py reads it for roofline.traffic."""
    import json
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def num(v, unit):
        x = float(v.replace(',', ''))
        return x * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3}.get(unit, 1.0)

    best = {}
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        key = next((k for k in ('presmooth', 'postsmooth', 'apply_p', 'update_xr') if 'k_' + k in name), None)
        if not key:
            continue
        g = r[col['Grid Size']]
        gsz = 1
        for t in g.strip('()').split(','):
            gsz *= int(t)
        rd, wr = col['dram__bytes_read.sum'], col['dram__bytes_write.sum']
        rec = {'dram_bytes': num(r[rd], units[rd]) + num(r[wr], units[wr]),
               'dram_read': num(r[rd], units[rd]), 'dram_write': num(r[wr], units[wr]),
               'time_us': num(r[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']]),
               'grid': g, 'kernel': name.split('(')[0]}
        if key not in best or gsz > best[key][0]:
            best[key] = (gsz, rec)
    json.dump({'source': path, 'kernels': {k: v[1] for k, v in best.items()}}, open(out_json, 'w'), indent=1)
    print(json.dumps({k: v[1] for k, v in best.items()}, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == '--traffic':
        traffic(sys.argv[1], sys.argv[3])
    else:
        raw(sys.argv[1])
        if len(sys.argv) > 2:
            sass(sys.argv[1])
