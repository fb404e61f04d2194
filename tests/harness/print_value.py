"""Prints value / ms_per_step (and any further keys named on the command line) of the JSON line in a bench.py output file.
   python tests/harness/print_value.py LABEL FILE [key ...]"""
import json
import sys

label, path, keys = sys.argv[1], sys.argv[2], sys.argv[3:]
for line in open(path):
    if line.startswith("{"):
        d = json.loads(line)
        print(label, d.get("value"), d.get("unit"), d.get("ms_per_step"), "ms")
        for k in keys:
            print(" ", k, json.dumps(d.get(k)))
