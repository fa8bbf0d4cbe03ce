"""Compact view of a bench.py JSON line:  python tools/show_bench.py <file> [keys...]"""
import json
import sys

line = [l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1]
d = json.loads(line)


def show(x, ind=0, maxlen=150):
    for k, v in x.items():
        if isinstance(v, dict):
            print(" " * ind + k + ":")
            show(v, ind + 2, maxlen)
        else:
            s = json.dumps(v)
            print(" " * ind + f"{k}: {s[:maxlen]}")


keys = sys.argv[2:]
show({k: d[k] for k in keys if k in d} if keys else d)
