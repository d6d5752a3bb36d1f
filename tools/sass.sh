#!/bin/bash
# rebuild the library + workload programs, disassemble gpupreagg_main of each cached cubin
cd /root/repo
rm -f pg_strom_b200/_cubin_cache/*.cubin
python -c "
from pg_strom_b200 import build as b; b.build_library(verbose=True)
from pg_strom_b200 import workloads; workloads.prebuild_programs()" 2>&1 | grep -v "warning #128\|for (cl_ulong\|\^\|^$\|Remark" | tail -30
cd pg_strom_b200/_cubin_cache
for f in *.cubin; do cuobjdump -sass -fun gpupreagg_main $f | grep -v "^\s*/\* 0x" | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-110 > /tmp/$f.sass; echo $f insts=$(grep -c "^\s*/\*[0-9a-f]*\*/" /tmp/$f.sass) ATOMS=$(grep -c ATOMS /tmp/$f.sass) DADD=$(grep -c "DADD" /tmp/$f.sass) $(cuobjdump -res-usage $f 2>/dev/null | grep -A1 "Function gpupreagg_main:" | tail -1 | cut -c1-40); done
