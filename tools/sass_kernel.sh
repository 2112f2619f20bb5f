#!/bin/bash
# SASS of one kernel of an object / library, one instruction per line:  tools/sass_kernel.sh FILE MANGLED_SUBSTRING
F=$1; K=$2
N=$(cuobjdump -elf $F 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*${K}[A-Za-z0-9_]*" | sort -u | head -1)
cuobjdump -sass -fun "$N" $F | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\*[0-9a-f]*\*/##; s#/\* 0x[0-9a-f]* \*/##; s/  */ /g'
