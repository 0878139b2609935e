#!/bin/bash
# registers / spills per kernel of one .cu file:  tools/ptxas_report.sh cocodr_b200/csrc/gemm.cu
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -c "$1" -o /tmp/_ptxas.o > /tmp/_ptxas.log 2>&1
grep -E "error" /tmp/_ptxas.log | head -5
python - <<'PY'
import re
t=open('/tmp/_ptxas.log').read()
for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s*(.*?spill loads)\n.*?Used (\d+) registers", t):
    name=m.group(1)
    import subprocess
    d=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
    d=re.sub(r'\(.*','',d).replace('void ','').replace('cdr::','').replace('(int)','').replace('(bool)','')
    print(f"{d:50s} regs={m.group(3):>3s}  {m.group(2)}")
PY
