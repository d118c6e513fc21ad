"""stdin: cuobjdump -sass output; stdout: one row per hot kernel with counts of the instructions that matter"""
import collections
import re
import subprocess
import sys

OPS = ('UTMALDG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'BAR.SYNC', 'DFMA', 'DADD', 'DMUL', 'LDG.E.128', 'STG.E.128', 'LDS.128', 'STS.128')
HOT = ('k_pred_tma', 'k_corr_tma', 'k_fft_lines_bs', 'k_fft_solve_bs', 'k_bulk_rows', 'k_fft_x_c2r_d', 'k_fft_x_r2c_v',
       'k_fft_lines_io', 'k_fft_solve_r<512', 'k_fft_lines_r<512', 'k_thomas_lp')
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for op in OPS:
            if op in line:
                cnt[cur][op] += 1
names = list(cnt)
dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.split('\n')
print('%-90s %s' % ('kernel', ' '.join('%9s' % o for o in OPS)))
for n, d in zip(names, dem):
    if any(k in d for k in HOT):
        print('%-90s %s' % (d[:90], ' '.join('%9d' % cnt[n][o] for o in OPS)))
