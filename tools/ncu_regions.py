"""Groups the SASS of a kernel in an .ncu-rep into regions of equal execution count (developer tool)."""
import csv
import subprocess
import sys
from collections import Counter


def main(path, rows_processed):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    idxs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    body = rows[idxs[0] + 2: idxs[1] if len(idxs) > 1 else None]
    n = float(rows_processed)
    tot = sum(int(r[5]) for r in body if len(r) > 5 and r[5].isdigit())
    print('kernel:', rows[idxs[0]][1][:90])
    print('warp instructions per row: %.1f' % (tot / n))
    cur, acc = None, []
    for r in body:
        if len(r) < 6 or not r[5].isdigit():
            continue
        c = int(r[5])
        m = c / n
        if cur is None or abs(m - cur[0]) > 0.08 * max(m, cur[0], 0.1):
            if cur:
                acc.append(cur)
            cur = [m, 0, 0, Counter()]
        cur[1] += 1
        cur[2] += c
        parts = r[1].strip().split()
        cur[3][parts[1] if parts[0].startswith('@') else parts[0]] += 1
    acc.append(cur)
    for m, cnt, c, ops in acc:
        if c / n > 5:
            print('x%.2f  static %3d  dyn/row %6.1f  %s' % (m, cnt, c / n, ops.most_common(7)))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
