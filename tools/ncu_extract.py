"""profiles/ from an ncu report of the contact kernels.

    ncu -i gpurun_out/prof_contact_<tag>.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_extract.py raw.csv

writes profiles/r01_ncu_full_contact_10M.csv (selected metrics, one column per
launch: k_neighbours, k_list_sort, k_slots) and profiles/r01_traffic.json (the
DRAM bytes and pipe utilisations bench.py quotes in `roofline`)."""
import csv
import json
import sys

KEEP = ('Kernel Name', 'gpu__time_duration', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg',
        'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate',
        'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct', 'launch__',
        'sm__inst_executed_pipe_fp64.avg', 'sm__pipe_fp64_cycles_active.avg',
        'sm__issue_active.avg', 'sm__warps_active.avg', 'sm__throughput.avg',
        'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed',
        'smsp__average_warps_issue_stalled', 'lts__t_sectors.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local')
LABELS = ['k_neighbours (list rebuild, about every 6th step)',
          'k_list_sort (list rebuild)', 'k_slots (every step)']


def main(raw, source):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open('profiles/r01_ncu_full_contact_10M.csv', 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + ['launch%d' % i
                                         for i in range(len(data))])
        for i, h in enumerate(hdr):
            if h.startswith(KEEP):
                w.writerow([h, units[i]] + [r[i] for r in data])

    def val(name, k, scale=None):
        i = hdr.index(name)
        v = float(data[k][i])
        if scale:
            v *= scale[units[i]]
        return v
    out = {'source': source,
           'config': {'bodies': 100000, 'particles': 10000000,
                      'skin_factor': 0.05}, 'kernels': {}}
    byte = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.}
    ms = {'ms': 1., 'us': 1e-3, 'ns': 1e-6, 's': 1e3}
    for k, lab in enumerate(LABELS):
        out['kernels'][lab] = {
            'dram_read_bytes': val('dram__bytes_read.sum', k, byte),
            'dram_write_bytes': val('dram__bytes_write.sum', k, byte),
            'ms': val('gpu__time_duration.sum', k, ms),
            'fp64_pipe_pct': val('sm__inst_executed_pipe_fp64.avg.'
                                 'pct_of_peak_sustained_active', k),
            'issue_active_pct': val('sm__issue_active.avg.'
                                    'pct_of_peak_sustained_elapsed', k),
            'threads_per_inst': val('smsp__thread_inst_executed_per_inst_'
                                    'executed.ratio', k),
            'registers': val('launch__registers_per_thread', k),
            'warp_instructions': val('smsp__inst_executed.sum', k)}
    kn, kl, ks = (out['kernels'][lab] for lab in LABELS)
    out['contact_dram_bytes_per_evaluation'] = \
        ks['dram_read_bytes'] + ks['dram_write_bytes']
    out['contact_dram_bytes_per_rebuild'] = sum(
        k[n] for k in (kn, kl) for n in ('dram_read_bytes',
                                         'dram_write_bytes'))
    out['note'] = ('per evaluation without list rebuild = k_slots alone; a '
                   'rebuild adds k_neighbours + k_list_sort '
                   '(contact_dram_bytes_per_rebuild)')
    json.dump(out, open('profiles/r01_traffic.json', 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else
         'ncu --set full --clock-control none --import-source on -k '
         'regex:k_slots|k_list_sort|k_neighbours --launch-skip 910 -c 3, '
         'python tools/ktime.py 100000 300 x 0.05: one list rebuild followed '
         'by one pair evaluation on the 10M pile after 300 settle steps')
