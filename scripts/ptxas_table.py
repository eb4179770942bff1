"""Registers / spills per score_row_kernel instantiation from the ptxas -v logs of build dirs (dev tool)."""
import re
import sys

for d in sys.argv[1:]:
    print(d)
    for f in ['k_row_whole', 'k_row_seg', 'k_row_sub']:
        t = open(f'{d}/{f}.log').read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", t):
            mm = re.search(r"score_row_kernelILi(\d+)ELi(\d+)ELi(\d+)ELb(\d)", m.group(1))
            if mm:
                print('  Q=%s SEG=%s MODE=%s DUMP=%s regs=%s spill=%s/%s' % (*mm.groups(), m.group(5), m.group(3), m.group(4)))
