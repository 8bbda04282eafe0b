"""Extract per-kernel DRAM traffic + duration from an ncu report into a small JSON that bench.py
reads for `roofline.traffic`.  usage: python profiles/ncu_traffic.py report.ncu-rep out.json workload [note]"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3,
        "usecond": 1.0, "msecond": 1e3}
NAMES = {"lookup_sort_kernel": "lookup", "lookup_gather_kernel": "lookup", "lookup_pool_kernel": "lookup", "update_tiles_kernel": "update",
         "update_fixup_kernel": "update_fixup", "interaction_fwd_kernel": "interaction_fwd",
         "interaction_bwd_kernel": "interaction_bwd",
         "interaction_fwd_mma_kernel": "interaction_fwd", "interaction_fwd_mma_ksplit_kernel": "interaction_fwd",
         "interaction_fwd_warp_kernel": "interaction_fwd", "interaction_bwd_warp_kernel": "interaction_bwd", "sort_small_kernel": "sort",
         "interaction_bwd_ring_kernel": "interaction_bwd", "interaction_bwd_ring2_kernel": "interaction_bwd"}


def main(path, out, workload, note=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        return float(r[col[key]].replace(",", "")) * UNIT.get(units[col[key]], 1.0)

    res = {}
    for r in rows[2:]:
        kname = r[col["Kernel Name"]]
        short = next((v for k, v in NAMES.items() if k in kname), None)
        if short is None:
            continue
        res[short] = {   # last launch of each kernel wins
            "kernel": kname.split("(")[0],
            "dram_bytes_read": val(r, "dram__bytes_read.sum"),
            "dram_bytes_write": val(r, "dram__bytes_write.sum"),
            "duration_us_under_ncu": val(r, "gpu__time_duration.sum"),
        }
    doc = {"workload": workload, "report": path.split("/")[-1], "kernels": res}
    if note:
        doc["note"] = note
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
