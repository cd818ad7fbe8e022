#!/bin/bash
# SASS of the hot kernels of the built library -> profiles/sass_*_<round>.txt (build container; no GPU needed)
set -e
cd "$(dirname "$0")/.."
R=${1:-r02}
so=apples_b200/libapples_b200.so
dump() { cuobjdump -sass -fun "$1" $so | sed -n '/Function :/,$p' | sed 's/[[:space:]]*\/\* 0x[0-9a-f]* \*\/$//' | grep -v "^\s*/\* 0x" > profiles/sass_$2_$R.txt; echo "$2: $(grep -c '/\*[0-9a-f]\{4\}\*/' profiles/sass_$2_$R.txt) instructions"; }
dump _Z15dense_tc_kernel6TcArgs dense_tc
dump _Z17select_nuc_kernelILb0EEv10SelectArgs select
dump _Z17place_smem_kernelILi0ELi128EEv9PlaceArgs place_smem128
dump "$(cuobjdump -sass $so | grep -o '_Z[0-9]*dense_aa_kernel[A-Za-z0-9_]*' | head -1)" dense_aa
dump "$(cuobjdump -sass $so | grep -o '_Z[0-9]*dense_nuc_kernelILb0[A-Za-z0-9_]*' | head -1)" dense_nuc
echo "tensor-core / TMA mnemonics in dense_tc:"; grep -o "UTCIMMA[A-Z0-9_.]*\|UBLKCP[A-Z0-9_.]*\|UTCBAR[A-Z0-9_.]*\|LDTM[A-Z0-9_.]*\|SYNCS[A-Z0-9_.]*" profiles/sass_dense_tc_$R.txt | sort | uniq -c
