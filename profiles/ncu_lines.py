"""Aggregate the SASS rows of `ncu --page source --print-source cuda,sass` per CUDA source line:
share of stall samples, share of executed warp instructions, average active threads.

usage: python profiles/ncu_lines.py report.ncu-rep [top_n]     (the capture needs --import-source on, -lineinfo)
"""
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                         capture_output=True, text=True).stdout
    kernels, seen, cur, fname, line = [], set(), None, None, None
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] in ('File Path', 'File Name'):
            fname = r[1].split('/')[-1]
            if cur is None or fname in seen:  # a file shows up again: the next kernel's listing starts
                cur, seen = {'agg': {}, 'src': {}}, set()
                kernels.append(cur)
            seen.add(fname)
            continue
        if cur is None or not r or not (r[0] == '' or r[0].isdigit()):
            continue
        if r[0] != '':
            line = (fname, int(r[0]))
            cur['src'][line] = r[1]
            if len(r) < 8:
                continue
        if len(r) > 10 and r[2].startswith('0x'):
            try:
                smp, ins, thr = int(r[6]), int(r[7]), int(r[8])
            except ValueError:
                continue
            a = cur['agg'].setdefault(line, [0, 0, 0])
            a[0] += smp
            a[1] += ins
            a[2] += thr
    for i, k in enumerate(kernels):
        agg = k['agg']
        tot = [sum(a[j] for a in agg.values()) for j in range(3)]
        if tot[1] == 0:
            continue
        print('===== kernel %d: %d stall samples, %d warp instructions, %.1f threads active on average'
              % (i, tot[0], tot[1], tot[2] / tot[1]))
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            print('%5.1f%% smp %5.1f%% inst thr %4.1f  %s:%d  %s' % (100 * a[0] / max(tot[0], 1), 100 * a[1] / tot[1],
                                                                   a[2] / max(a[1], 1), key[0], key[1],
                                                                   k['src'][key].strip()[:90]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
