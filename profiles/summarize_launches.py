"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [top]"""
import collections
import csv
import re
import sys


def short(n):
    m = re.search(r'(\w+_kernel|\w+Kernel\w*|elementwise\w*)', n)
    base = m.group(1) if m else n[:40]
    if 'gemm_nn_kernel' in n or 'gemm_tn_kernel' in n:
        t = re.search(r'<(.*?)>\(', n)
        if t:
            base += '<' + t.group(1).replace('tmk::', '').replace('(bool)', '').replace('(int)', '') + '>'
    return base


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', '')) / 1000.0
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"{'kernel':92s} {'n':>4s} {'total us':>9s} {'avg us':>8s} {'share':>6s}")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{k[:92]:92s} {c:4d} {t:9.1f} {t / c:8.1f} {100 * t / tot:5.1f}%")
    print(f"total {tot:.1f} us over {len(data)} launches")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
