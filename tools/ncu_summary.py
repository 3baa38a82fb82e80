#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel, SASS opcode mix and stall reasons.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import collections, csv, io, re, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    extra = ['--kernel-name', 'regex:' + kre] if kre else []
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv'] + extra))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('== kernel:', r[hdr.index('Kernel Name')][:110])
        for k in KEYS:
            if k in hdr:
                print('   %-75s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
    src = ncu(['-i', rep, '--page', 'source', '--csv'] + extra)
    rows = list(csv.reader(io.StringIO(src)))
    # the source page may hold several kernels; take the first block
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
    hdr = rows[hi]
    ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    ops, samp, st, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] == 'Address':
            break
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ia].strip())
        op = m.group(2) if m else r[ia].strip()
        n = int(r[ie]); ops[op] += n; tot += n; samp[op] += int(r[isamp])
        for i in stall:
            st[hdr[i]] += int(r[i] or 0)
    print('-- SASS opcode mix (warp instructions executed), first kernel; total', tot)
    for op, n in ops.most_common(18):
        print('   %-10s %12d %5.1f%%   samples %d' % (op, n, 100.0 * n / tot, samp[op]))
    s = sum(st.values())
    print('-- warp stall samples')
    for k, v in st.most_common(10):
        print('   %-26s %8d %5.1f%%' % (k, v, 100.0 * v / max(s, 1)))


if __name__ == '__main__':
    main()
