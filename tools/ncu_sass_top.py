#!/usr/bin/env python
"""Top SASS instructions of one kernel of an .ncu-rep by stall samples, with their dominant stall reasons and the CUDA line they come from:
python tools/ncu_sass_top.py report.ncu-rep <kernel regex> [top]"""
import csv
import subprocess
import sys


def main(path, regex, top=30):
    text = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + regex,
                           '--launch-skip', '0', '--launch-count', '1'], capture_output=True, text=True).stdout
    cur, hdr, line, out = None, None, None, []
    for r in csv.reader(text.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            isamp = hdr.index('# Samples')
            stall_cols = [(i, h[len('stall_'):]) for i, h in enumerate(hdr) if h.startswith('stall_') and not h.endswith('_not_issued')]
            continue
        if hdr is None:
            continue
        if r[0]:
            line = '{}:{}'.format(cur, r[0])
            continue
        try:
            n = int(r[isamp])
        except (ValueError, IndexError):
            continue
        if n:
            reasons = sorted(((int(r[i] or 0), name) for i, name in stall_cols if i < len(r) and (r[i] or '0').isdigit()), reverse=True)[:2]
            out.append((n, line, r[3].strip(), ', '.join('{} {}'.format(nm, v) for v, nm in reasons if v)))
    tot = sum(o[0] for o in out) or 1
    for o in sorted(out, reverse=True)[:top]:
        print('{:5.2f}%  {:<28s} {:<60s} {}'.format(100 * o[0] / tot, o[1], o[2][:60], o[3]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
