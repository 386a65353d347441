"""profiles/roofline_kernel_traffic.json <- dram__bytes_read.sum + dram__bytes_write.sum of the roofline kernel
(first matching launch without the residual segment) in an `ncu --set full` capture.
    python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep umma_conv3p_kernel"""
import csv, json, os, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
ki = h.index("Kernel Name")
ri, wi, ti = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
tp = h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
cands = [r for r in data if pat in r[ki]]
best = min(cands, key=lambda r: float(r[ri]) * scale[units[ri]])   # the launch that reads only x (no residual operand)
rd, wr = float(best[ri]) * scale[units[ri]], float(best[wi]) * scale[units[wi]]
d = {"kernel": best[ki].split("(")[0], "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes_per_launch": rd + wr,
     "duration_us_under_ncu": float(best[ti]), "tensor_pipe_active_pct": float(best[tp]),
     "algorithmic_bytes": 128 * 32 * 32 * 128 * 2 * 2 + 9 * 128 * 128 * 2,
     "note": "in-step launch (L2 holds the output: DRAM writes are deferred to later kernels), ncu --set full --clock-control none",
     "source": os.path.basename(rep)}
json.dump(d, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_kernel_traffic.json"), "w"), indent=1)
print(d)
