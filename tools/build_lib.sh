#!/bin/bash
# Build libtopicgcn.so in-tree (sm_100a) and print the register / spill summary of the role kernels.
set -e
cd "$(dirname "$0")/.."
python -c "
import sys; sys.path.insert(0,'graph-convolutional-networks-for-text-classification_b200')
import build; build.build()
" 2>&1 | grep -v "ptxas info" | tail -4
python - <<'PY'
import re,subprocess
txt=open('graph-convolutional-networks-for-text-classification_b200/build/tg_roles2.cu.ptxas.log').read()
blocks=re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
for name,stack,ss,sl,regs in blocks:
    if 'roles2' in name and (int(ss) or int(sl) or 'roles2_kernel' in name):
        dem=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()[:80]
        print(regs,'regs, spill',ss,sl,dem)
PY
