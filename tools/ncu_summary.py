"""Summarise an .ncu-rep (raw page, base units) into a compact per-launch CSV + markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_conv.ncu-rep profiles/r01_conv_tc_ncu
"""
import csv
import io
import subprocess
import sys

COLS = [
    ('gpu__time_duration.sum', 'dur_us', 1e-3),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct', 1),
    # the tcgen05 (UTCHMMA) work itself: executed bf16 tensor math ops (2 per MAC), the HMMA-subpipe busy fraction, the
    # warp-level tensor instructions and the shared-memory operand wavefronts the MMAs read
    ('sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum', 'tensor_ops_G', 1e-9),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'hmma_subpipe_pct', 1),
    ('sm__inst_executed_pipe_tensor_subpipe_hmma.sum', 'utchmma_inst', 1),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum', 'smem_wf_A_M', 1e-6),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum', 'smem_wf_B_M', 1e-6),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wf_lsu_M', 1e-6),
    ('sm__cycles_elapsed.max', 'sm_cycles', 1),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct', 1),
    ('dram__bytes_read.sum', 'dram_rd_MB', 1e-6),
    ('dram__bytes_write.sum', 'dram_wr_MB', 1e-6),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct', 1),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'l2_to_sm_MB', 1e-6),
    ('l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed', 'l2_to_sm_pct', 1),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts_pct', 1),
    ('smsp__inst_executed.sum', 'warp_inst_M', 1e-6),
    ('sm__inst_executed.avg.per_cycle_elapsed', 'ipc', 1),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct', 1),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu_pipe_pct', 1),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conflicts_M', 1e-6),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_sb', 1),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall_short_sb', 1),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall_not_selected', 1),
    ('launch__registers_per_thread', 'regs', 1),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct', 1),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--print-units', 'base'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = {}
    for name, short, _ in COLS:
        for i, h in enumerate(hdr):
            if h == name or h.endswith('.' + name):
                idx[short] = i
                break
    k_name, k_grid, k_block = hdr.index('Kernel Name'), hdr.index('Grid Size'), hdr.index('Block Size')
    table = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[k_name]
        name = name.split('(')[0].replace('void ', '').replace('rcu::', '')
        rec = {'kernel': name, 'grid': r[k_grid].replace(' ', ''), 'block': r[k_block].replace(' ', '')}
        for mname, short, scale in COLS:
            if short in idx:
                try:
                    rec[short] = round(float(r[idx[short]].replace(',', '')) * scale, 3)
                except ValueError:
                    rec[short] = None
        table.append(rec)
    keys = ['kernel', 'grid', 'block'] + [s for _, s, _ in COLS if s in idx]
    with open(out + '.csv', 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for rec in table:
            w.writerow(rec)
    with open(out + '.md', 'w') as f:
        f.write('| # | ' + ' | '.join(keys) + ' |\n|' + '---|' * (len(keys) + 1) + '\n')
        for i, rec in enumerate(table):
            f.write('| %d | ' % i + ' | '.join(str(rec.get(k, '')) for k in keys) + ' |\n')
    print('wrote', out + '.csv', out + '.md', len(table), 'launches')


if __name__ == '__main__':
    main()
