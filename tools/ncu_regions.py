"""Where do the executed instructions and stall samples of a kernel fall?  (ncu --page source --csv, split into SASS regions
of similar execution count)   usage: ncu_regions.py report.ncu-rep kernel-regex [instance-from-end]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
inst = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
st = starts[-inst]
end = starts[-inst + 1] if inst > 1 else len(rows)
hdr = rows[st + 1]
ia, isrc, ismp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
data = [(r[isrc].strip(), int(r[ia] or 0), int(r[ismp] or 0)) for r in rows[st + 2:end] if len(r) > ia]
tot, ts = sum(d[1] for d in data), max(1, sum(d[2] for d in data))
print(rows[st][1][:80], 'total warp-inst', tot, 'sass', len(data))
i = 0
while i < len(data):
    j, base = i, data[i][1]
    while j < len(data) and abs(data[j][1] - base) <= 0.25 * max(base, 1):
        j += 1
    s, sm = sum(d[1] for d in data[i:j]), sum(d[2] for d in data[i:j])
    if s > 0.004 * tot or sm > 0.01 * ts:
        print('%5d-%5d n=%4d  exec/inst=%9d  %5.1f%% inst %5.1f%% samples  %s' % (i, j - 1, j - i, base, 100 * s / tot, 100 * sm / ts, data[i][0][:45]))
    i = j
if len(sys.argv) > 5:
    for k in range(int(sys.argv[4]), int(sys.argv[5])):
        print(k, data[k][1], data[k][2], data[k][0][:90])
