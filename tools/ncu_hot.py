"""Summarise an ncu report: headline metrics per launch and the hottest SASS lines with their context (development aid).
usage: python tools/ncu_hot.py report.ncu-rep [n_top] [context]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} [{units[i]}]:", [r[i] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {'hdr': None, 'rows': []}; secs.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = r; continue
    cur['rows'].append(r)
seen = set()
for si, sec in enumerate(secs):
    h = sec['hdr']; isrc = h.index('Source'); isamp = h.index('# Samples')
    tot = sum(int(r[isamp] or 0) for r in sec['rows'])
    key = (tot, len(sec['rows']))
    if key in seen: continue          # ncu repeats each kernel section
    seen.add(key)
    order = sorted(range(len(sec['rows'])), key=lambda i: -int(sec['rows'][i][isamp] or 0))[:ntop]
    print(f"== section {si}: {tot} samples")
    for idx in order:
        print(f"  -- instr {idx}: {100 * int(sec['rows'][idx][isamp]) / tot:.1f}%")
        for j in range(max(0, idx - ctx), idx + 1):
            print(f"     {j:5d} {sec['rows'][j][isamp]:>6}  {sec['rows'][j][isrc][:100]}")
