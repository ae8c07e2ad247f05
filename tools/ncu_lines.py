#!/usr/bin/env python
"""Stall samples and executed instructions per CUDA source line of one kernel of an .ncu-rep (needs --import-source on):
python tools/ncu_lines.py report.ncu-rep <kernel regex> [top]"""
import csv
import subprocess
import sys


def main(path, regex, top=40):
    text = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + regex,
                           '--launch-skip', '0', '--launch-count', '1'], capture_output=True, text=True).stdout
    cur, hdr, out = None, None, []
    for r in csv.reader(text.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            isamp, iinst = hdr.index('# Samples'), hdr.index('Instructions Executed')
            continue
        if hdr is None or not r[0]:
            continue      # SASS rows (empty line number) are already summed in their source line's row
        try:
            out.append((int(r[isamp]), int(r[iinst]), cur, r[0], r[1].strip()[:120]))
        except (ValueError, IndexError):
            pass
    tot, toti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    print('total samples {}, warp instructions {}'.format(tot, toti))
    for o in sorted(out, reverse=True)[:top]:
        print('{:5.1f}% samples {:5.1f}% inst  {}:{}  {}'.format(100 * o[0] / tot, 100 * o[1] / toti, o[2], o[3], o[4]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
