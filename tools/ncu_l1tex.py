"""Per-kernel l1tex / tensor / issue breakdown from an .ncu-rep (wavefronts per SM; diagnostic)."""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
idx = {n: i for i, n in enumerate(h)}
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg', 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg',
        'smsp__inst_executed.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__issue_active.avg.per_cycle_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for d in rows[2:]:
    print('==', d[idx['Kernel Name']][:100])
    sms = 148.0
    for w in want:
        if w in idx:
            extra = ''
            if w.endswith('.sum') and 'wavefronts' in w:
                extra = f"  (per SM {float(d[idx[w]].replace(',', '')) / sms:.0f})"
            print('   ', w, u[idx[w]], d[idx[w]], extra)
