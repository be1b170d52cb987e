# usage: bash scripts/ptxas_stats.sh "<extra nvcc flags>" [kernel-regex]   — registers / spills of the tracing kernels
F="$1"; K="${2:-k_vps_pool|k_twostates_pool}"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -Xptxas -v $F -c -o /tmp/rcv_trace_stats.o \
  "$(dirname "$0")/../recursive_control_variates_b200/csrc/rcv_trace.cu" 2>&1 | python3 -c "
import re, sys
name = None
for l in sys.stdin:
    m = re.search(r\"Compiling entry function '(\S+)'\", l)
    if m: name = m.group(1); spill = ''
    elif 'spill' in l: spill = l.strip()
    elif 'Used' in l and name and re.search(r'$K', name):
        print(name[8:60], '|', re.search(r'Used \d+ registers', l).group(0), '|', spill)
"
