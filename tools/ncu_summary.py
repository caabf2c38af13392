"""Summarise an .ncu-rep: key metrics per kernel + the hottest SASS instructions (needs ncu on PATH)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform", "l1tex__data_bank_reads.sum", "l1tex__data_bank_writes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__occupancy_limit", "smsp__cycles_active.avg", "l1tex__throughput.avg.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "launch__shared_mem_per_block_dynamic"]
idx = {h: i for i, h in enumerate(hdr)}
for r in data:
    print("=== ", r[idx["Kernel Name"]][:110])
    for h in hdr:
        if any(h.startswith(w) for w in want) and "per_second" not in h and "peak_sustained_elapsed" not in h.replace("lts__throughput.avg.pct_of_peak_sustained_elapsed", "").replace("sm__throughput.avg.pct_of_peak_sustained_elapsed", ""):
            print(f"   {h} [{units[idx[h]]}] = {r[idx[h]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
for sec in src.split('"Kernel Name",')[1:]:
    lines = sec.split("\n")
    name = lines[0]
    rdr = csv.reader(io.StringIO("\n".join(lines[1:])))
    h = next(rdr)
    rs = [r for r in rdr if len(r) == len(h)]
    si, ie = h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[si]) for r in rs)
    print("=== hot SASS of", name[:100], "total samples", tot)
    ranked = sorted(enumerate(rs), key=lambda t: -int(t[1][si]))[:topn]
    for i, r in sorted(ranked):
        st = {h[c].replace("stall_", ""): int(r[c]) for c in stall_cols if int(r[c]) > 0}
        print(f"   {i:4d} {r[1].strip()[:64]:64s} smp {r[si]:>5s} exec {r[ie]:>7s} {st}")
