"""profiles/traffic_feas_kernel.json from an ncu --set full capture of the default bench config -- development tool.
usage: ncu_traffic.py rep kernel_name n_d n_theta "source note" """
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import source_hash  # noqa: E402

rep, kernel, n_d, n_t, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}


def val(name, scale={"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}):
    v, u = d[name]
    return float(v.replace(",", "")) * scale.get(u, 1)


out = {"source": note, "kernel": kernel, "n_d": n_d, "n_theta": n_t,
       "dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")),
       "gpu_time_ms": val("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6}.get(d["gpu__time_duration.sum"][1], 1),
       "fp64_pipe_active_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "shared_wavefronts_pct": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
       "registers_per_thread": int(val("launch__registers_per_thread")),
       "warp_instructions": int(val("smsp__inst_executed.sum")),
       # bench.py quotes these figures only while the kernel sources hash to this value
       "source_sha256": source_hash()}
json.dump(out, open("profiles/traffic_feas_kernel.json", "w"), indent=1)
print(out)
