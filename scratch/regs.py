#!/usr/bin/env python
"""List registers / spills / smem of kernels from an `nvcc -Xptxas -v` log (default /tmp/build.log)."""
import re, subprocess, sys
txt = open(sys.argv[1] if len(sys.argv) > 1 else '/tmp/build.log').read()
pat = sys.argv[2] if len(sys.argv) > 2 else 'stream'
for b in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    if pat not in dem: continue
    m = re.search(r"Used (\d+) registers", b); sp = re.search(r"(\d+) bytes spill stores", b); st = re.search(r"(\d+) bytes stack frame", b)
    print('%-110s regs %3s spill %4s stack %4s' % (dem.replace('pd::ts::','').replace('(pd::WarpParams, pd::ts::StreamCfg)','')[:110], m.group(1) if m else '?', sp.group(1) if sp else '?', st.group(1) if st else '?'))
