"""Summarise an .ncu-rep (one kernel launch): key metrics + stall samples by reason and by SASS opcode.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sass__inst_executed_register_spilling',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = []
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out.append('section,name,unit,value')
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            out.append(f'metric,{h},{u},{v}')
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv']))))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    samp, execd = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']].strip())
        op = m.group(2).split('.')[0] if m else '?'
        samp[op] += int(r[idx['# Samples']] or 0)
        execd[op] += int(r[idx['Instructions Executed']] or 0)
    tot = sum(samp.values()) or 1
    for op, s in samp.most_common(24):
        out.append(f'opcode,{op},samples_pct/exec_M,{100 * s / tot:.1f}/{execd[op] / 1e6:.1f}')
    text = '\n'.join(out) + '\n'
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text)
    print(text)


if __name__ == '__main__':
    main()
