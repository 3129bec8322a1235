"""profiles/r02_traffic.json from the round's own `ncu --set full` capture of the headline step.

    ncu --set full --clock-control none -k regex:'gemm_tcgen05|seg_reduce' -c 6 -o gpurun_out/r02_layer python bench.py --steps 2 --warmup 3 --skip-e2e ...
    python tools/ncu_traffic.py gpurun_out/r02_layer.ncu-rep profiles/r02_traffic.json

Sums dram__bytes_read.sum + dram__bytes_write.sum over the kernels of ONE RGCN layer (one gemm_tcgen05_kernel + one
seg_reduce*_kernel launch: the last pair in the capture) -- the `roofline.traffic` of bench.py."""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


kernels = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    l2 = to_bytes(r[col["l1tex__m_xbar2l1tex_read_bytes.sum"]], units[col["l1tex__m_xbar2l1tex_read_bytes.sum"]]) if "l1tex__m_xbar2l1tex_read_bytes.sum" in col else None
    us = float(r[col["gpu__time_duration.sum"]])
    kernels.append({"kernel": name, "dram_read_bytes": rd, "dram_write_bytes": wr, "l2_to_sm_read_bytes": l2, "duration_us": us,
                    "tensor_pipe_active_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in col else None})
gemm = [k for k in kernels if "gemm_tcgen05" in k["kernel"]][-1]
seg = [k for k in kernels if "seg_reduce" in k["kernel"]][-1]
layer = gemm["dram_read_bytes"] + gemm["dram_write_bytes"] + seg["dram_read_bytes"] + seg["dram_write_bytes"]
json.dump({"source": rep, "what": "dram__bytes_read.sum + dram__bytes_write.sum of one RGCN layer (last gemm_tcgen05_kernel + seg_reduce kernel of the capture), ncu --set full --clock-control none",
           "dram_bytes_per_layer": layer, "kernels": [gemm, seg]}, open(out, "w"), indent=1)
print(json.dumps({"dram_bytes_per_layer": layer}))
