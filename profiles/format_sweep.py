"""Table view of a `bench.py --workload sweep` JSON line.  usage: python profiles/format_sweep.py sweep.json > sweep.txt"""
import json
import sys

r = json.load(open(sys.argv[1]))
s = r["summary"]
print(f"bench.py --workload sweep: one table, 2^20 lookups per launch (B = 2^20 / P), f32 rows; % = algorithmic bytes / time / "
      f"{s['hbm_peak_gbs']} GB/s ({s['peak_source']} HBM copy peak); algorithmic bytes per SURVEY 8(d): lookup L*D*4 + B*D*4 + L*4, "
      "update B*D*4 + L*8 + 2*U*D*4 (U = distinct rows)")
print(f"cases {s['cases']}; geometric mean {r['value']:.0f} GB/s = {100 * s['frac_hbm_geomean']:.0f}% of peak; D >= 64: "
      f"{100 * s['frac_hbm_geomean_D_ge_64']:.0f}%; cases >= 70%: {s['cases_at_or_above_70pct']}; cases < 50%: {s['cases_below_50pct']}")
print("      rows    D   P  zipf | lookup            | sort us  | update tiles+fixup | lookup+sort+update")
for c in r["sweep"]:
    print(f"{c['rows']:>10} {c['D']:>4} {c['P']:>3} {c['zipf']:>5} | {c['lookup_us']:8.1f}us {100 * c['lookup_frac']:4.0f}% | {c['sort_us']:8.1f} | "
          f"{c['update_us']:8.1f}us {100 * c['update_frac']:4.0f}% | {c['total_us']:8.1f}us {100 * c['frac_hbm']:4.0f}%")
