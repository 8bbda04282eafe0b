"""Pretty-print benchmarks/hotpath.py JSON lines.  usage: python profiles/show_hot.py file.json ..."""
import json
import sys

KEYS = ["lookup", "sort", "sort_plus_update", "update_only", "embedding_lookup_plus_update", "interaction_fwd", "interaction_bwd"]
for f in sys.argv[1:]:
    for line in open(f):
        d = json.loads(line)
        print(f"== {f}: {d['label']} tables={d['tables']} D={d['D']} B={d['B']} P={d['P']} zipf={d['zipf']} "
              f"distinct={d['distinct_rows_per_launch']:.0f}")
        for k in KEYS:
            if k in d:
                v = d[k]
                print(f"   {k:30s} {v['us']:9.2f} us  {v.get('gbs', 0):8.1f} GB/s  {100 * v.get('frac_hbm', 0):5.1f}% of HBM peak")
