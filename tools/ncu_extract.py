"""profiles/ from an ncu report of one list-rebuild step and one reuse step
(tools/prof_step.py under `ncu --set full --profile-from-start off`).

    ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_extract.py raw.csv r02

writes profiles/<tag>_ncu_full_step_10M.csv (selected metrics, one column per
launch, in launch order) and profiles/<tag>_traffic.json (the DRAM bytes and
utilisations bench.py quotes in `roofline`)."""
import csv
import json
import sys

KEEP = ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg',
        'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate',
        'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct',
        'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit',
        'sm__inst_executed_pipe_fp64.avg', 'sm__inst_executed_pipe_fma.avg',
        'sm__inst_executed_pipe_alu.avg', 'sm__inst_executed_pipe_xu.avg',
        'sm__inst_executed_pipe_lsu.avg',
        'sm__issue_active.avg', 'sm__warps_active.avg', 'sm__throughput.avg',
        'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed',
        'smsp__average_warps_issue_stalled', 'lts__t_sectors.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local')


def main(raw, tag):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open('profiles/%s_ncu_full_step_10M.csv' % tag, 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + ['launch%d' % i
                                         for i in range(len(data))])
        for i, h in enumerate(hdr):
            if h.startswith(KEEP):
                w.writerow([h, units[i]] + [r[i] for r in data])
    byte = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.}
    ms = {'ms': 1., 'us': 1e-3, 'ns': 1e-6, 's': 1e3}

    def val(r, name, scale=None):
        i = hdr.index(name)
        v = float(r[i])
        if scale:
            v *= scale[units[i]]
        return v
    # first occurrence of every kernel = the rebuild step; the second
    # k_filter / k_slots = a step that reuses the lists
    seen = {}
    for r in data:
        name = r[hdr.index('Kernel Name')]
        short = name.split('::')[-1].split('(')[0]
        if 'k_slots' in short:
            short = 'k_slots (compact)' if short.endswith('1>') else \
                'k_slots (all particles; returns at once)'
        seen.setdefault(short, []).append(r)
    out = {'source': 'gpurun_out/%s_prof.ncu-rep: ncu --set full '
           '--clock-control none --import-source on --profile-from-start off '
           'python tools/prof_step.py 100000 2000 0.075 (one list-rebuild '
           'step + one reuse step after 2000 settle steps)' % tag,
           'config': {'bodies': 100000, 'particles': 10000000,
                      'skin_factor': 0.075}, 'kernels': {}}
    for short, rs in seen.items():
        r = rs[-1] if short.startswith(('k_filter', 'k_slots')) else rs[0]
        out['kernels'][short] = {
            'ms': val(r, 'gpu__time_duration.sum', ms),
            'dram_read_bytes': val(r, 'dram__bytes_read.sum', byte),
            'dram_write_bytes': val(r, 'dram__bytes_write.sum', byte),
            'dram_pct_of_peak': val(r, 'gpu__dram_throughput.avg.'
                                    'pct_of_peak_sustained_elapsed'),
            'fp64_pipe_pct': val(r, 'sm__inst_executed_pipe_fp64.avg.'
                                 'pct_of_peak_sustained_active'),
            'issue_active_pct': val(r, 'sm__issue_active.avg.'
                                    'pct_of_peak_sustained_elapsed'),
            'threads_per_inst': val(r, 'smsp__thread_inst_executed_per_inst_'
                                    'executed.ratio'),
            'registers': int(val(r, 'launch__registers_per_thread')),
            'warp_instructions': val(r, 'smsp__inst_executed.sum')}
    k = out['kernels']
    out['contact_dram_bytes_per_evaluation'] = sum(
        k[n]['dram_read_bytes'] + k[n]['dram_write_bytes']
        for n in k if n.startswith(('k_filter', 'k_slots', 'k_sparse')))
    out['contact_dram_bytes_per_rebuild'] = sum(
        k[n]['dram_read_bytes'] + k[n]['dram_write_bytes']
        for n in k if n.startswith(('k_neighbours', 'k_list_sort')))
    with open('profiles/%s_traffic.json' % tag, 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'r02')
