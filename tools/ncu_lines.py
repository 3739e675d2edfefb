"""Per-source-line totals of an ncu report (built with -lineinfo, captured with
--import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_filter [launch] [top]

prints, for the lines of the kernel with the most executed warp instructions:
instructions executed, share, average active threads, stall samples."""
import csv
import io
import subprocess
import sys


def _num(v):
    # (lines without a count show '-' or an empty cell)
    try:
        return int(v)
    except ValueError:
        return 0


def lines(rep, kernel, launch=0):
    out = subprocess.run(
        ['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
         'cuda,sass', '--kernel-name', 'regex:' + kernel, '--launch-skip',
         str(launch), '--launch-count', '1'], capture_output=True,
        text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    res = {}
    fname, hdr = None, None
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif len(r) > 5 and r[0] == 'Line No':
            hdr = r
            ci = hdr.index('Instructions Executed')
            ct = hdr.index('Thread Instructions Executed')
            cs = hdr.index('# Samples')
        elif hdr and len(r) > 5 and r[0] != '':
            key = (fname, int(r[0]))
            e = res.setdefault(key, [r[1], 0, 0, 0])
            e[1] += _num(r[ci])
            e[2] += _num(r[ct])
            e[3] += _num(r[cs])
    return res


if __name__ == '__main__':
    rep, kernel = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    res = lines(rep, kernel, launch)
    tot = sum(e[1] for e in res.values()) or 1
    tots = sum(e[3] for e in res.values()) or 1
    print('%s: %d warp instructions, %d samples' % (kernel, tot, tots))
    for key, e in sorted(res.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-16s %5d %6.2f%% inst %5.1f thr %6.2f%% smp  %s' % (
            key[0][:16], key[1], 100. * e[1] / tot,
            e[2] / max(e[1], 1), 100. * e[3] / tots, e[0].strip()[:90]))
