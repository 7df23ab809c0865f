#!/bin/bash
# Compile libugf with -Xptxas -v and print one line per kernel: registers, spills, smem.
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off --shared -Xptxas -v \
  -o /tmp/libugf_ptxas.so unigasfoam_b200/csrc/ugf_api.cu 2>&1 | python3 -c "
import sys,re,subprocess
name=None
for l in sys.stdin:
    m=re.search(r\"Compiling entry function '(\S+)'\",l)
    if m: name=subprocess.run(['c++filt',m.group(1)],capture_output=True,text=True).stdout.strip().split('(')[0]; continue
    if 'spill' in l: sp=l.strip()
    m=re.search(r'Used (\d+) registers.*?(?:, (\d+) bytes smem)?$',l.strip())
    if m and name: print(f'{name:70s} regs={m.group(1):>3s} smem={m.group(2) or 0} | {sp}'); name=None
    if 'error' in l or 'warning' in l: print(l.rstrip())
"
